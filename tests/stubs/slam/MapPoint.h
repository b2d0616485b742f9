// TEST STUB of the slice of ORB_SLAM::MapPoint (reference include/MapPoint.h) the matcher facade reads.
#pragma once
#include <map>
#include <opencv2/core/core.hpp>
namespace ORB_SLAM {
class KeyFrame;
class MapPoint {
public:
    float GetMinDistanceInvariance() { return mfMinDistance; }   // MapPoint.h:79
    float GetMaxDistanceInvariance() { return mfMaxDistance; }   // MapPoint.h:80
    cv::Mat GetNormal() { return mNormal; }                      // MapPoint.h:46
    bool IsInKeyFrame(KeyFrame *kf) { return t_obs.count(kf) != 0; }                            // MapPoint.h:54
    int GetIndexInKeyFrame(KeyFrame *kf) { return t_obs.count(kf) ? (int)t_obs[kf] : -1; }      // MapPoint.h:53
    void Replace(MapPoint *p) { t_replaced_by = p; mbBad = true; }                               // MapPoint.h:62, MapPoint.cc:126-158
    void AddObservation(KeyFrame *kf, size_t idx) { t_obs[kf] = idx; }                          // MapPoint.h:50
    std::map<KeyFrame *, size_t> t_obs;   // test-only record of mObservations
    MapPoint *t_replaced_by;
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
    MapPoint() : t_replaced_by(NULL), mfMinDistance(0), mfMaxDistance(0), mTrackProjX(0), mTrackProjY(0), mbTrackInView(false), mnTrackScaleLevel(0), mTrackViewCos(0), mbBad(false) {}
};
}
