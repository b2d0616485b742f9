// oracle/ref_shim: shadows the reference's include/Converter.h (which pulls in g2o + Eigen, absent here).  Only
// toDescriptorVector is used on this path (Frame.cc:284, KeyFrame.cc:60); it is restated from src/Converter.cc:28-36.
#ifndef ORB_REF_SHIM_CONVERTER_H
#define ORB_REF_SHIM_CONVERTER_H
#include <vector>
#include <opencv2/core/core.hpp>
namespace ORB_SLAM {
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat &Descriptors) {
        std::vector<cv::Mat> vDesc;
        vDesc.reserve(Descriptors.rows);
        for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
        return vDesc;
    }
};
}  // namespace ORB_SLAM
#endif
