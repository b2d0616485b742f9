// oracle/ref_shim: boost::mutex / scoped_lock on std::mutex (the drivers are single-threaded; the locks are kept honest)
#ifndef ORB_REF_SHIM_BOOST_THREAD_HPP
#define ORB_REF_SHIM_BOOST_THREAD_HPP
#include <mutex>
namespace boost {
class mutex {
public:
    typedef std::unique_lock<std::recursive_mutex> scoped_lock_base;
    class scoped_lock {
    public:
        explicit scoped_lock(mutex &m) : l(m.m) {}
    private:
        std::unique_lock<std::recursive_mutex> l;
    };
    mutex() {}
    mutex(const mutex &) {}              // objects holding a mutex are copied by the reference (Frame copies, vectors)
    mutex &operator=(const mutex &) { return *this; }
private:
    std::recursive_mutex m;              // the reference re-locks the same mutex along some paths of KeyFrame.cc
};
}  // namespace boost
#endif
