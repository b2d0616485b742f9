// TEST STUB of the slice of ORB_SLAM::MapPoint (reference include/MapPoint.h) the matcher facade reads.
#pragma once
#include <opencv2/core/core.hpp>
namespace ORB_SLAM {
class KeyFrame;
class MapPoint {
public:
    float GetMinDistanceInvariance() { return mfMinDistance; }   // MapPoint.h:79
    float GetMaxDistanceInvariance() { return mfMaxDistance; }   // MapPoint.h:80
    cv::Mat GetNormal() { return mNormal; }                      // MapPoint.h:46
    bool IsInKeyFrame(KeyFrame *) { return false; }              // MapPoint.h:54
    int GetIndexInKeyFrame(KeyFrame *) { return -1; }            // MapPoint.h:53
    void Replace(MapPoint *) {}                                  // MapPoint.h:62
    void AddObservation(KeyFrame *, size_t) {}                   // MapPoint.h:50
    float mfMinDistance, mfMaxDistance;
    cv::Mat mNormal;
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
