// oracle/ref_shim: the reference includes <ros/ros.h> but only uses its assert / logging macros on this path
#ifndef ORB_REF_SHIM_ROS_H
#define ORB_REF_SHIM_ROS_H
#include <cstdio>
#include <cstdlib>
#define ROS_ASSERT(cond) do { if (!(cond)) { std::fprintf(stderr, "ROS_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); std::abort(); } } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#endif
