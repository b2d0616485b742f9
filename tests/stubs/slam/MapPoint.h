// TEST STUB of the slice of ORB_SLAM::MapPoint (reference include/MapPoint.h) the matcher facade reads.
#pragma once
#include <opencv2/core/core.hpp>
namespace ORB_SLAM {
class MapPoint {
public:
    cv::Mat GetWorldPos() { return mWorldPos; }          // 3x1 CV_32F (MapPoint.h:45)
    bool isBad() { return mbBad; }                       // MapPoint.h:60
    cv::Mat GetDescriptor() { return mDescriptor; }      // 1x32 CV_8U (MapPoint.h:70)
    // tracking variables filled by Frame::isInFrustum (MapPoint.h:85-91)
    float mTrackProjX, mTrackProjY;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    cv::Mat mWorldPos, mDescriptor;
    bool mbBad;
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mbTrackInView(false), mnTrackScaleLevel(0), mTrackViewCos(0), mbBad(false) {}
};
}
