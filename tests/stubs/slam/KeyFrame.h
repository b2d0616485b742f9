// TEST STUB of the slice of ORB_SLAM::KeyFrame (reference include/KeyFrame.h:44-140) that ORBmatcher calls.
// Exists only so that the facade's KeyFrame-level methods can be compile-checked without the SLAM sources.
#pragma once
#include <cstddef>
#include <map>
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include "MapPoint.h"
namespace DBoW2 {
class FeatureVector : public std::map<unsigned int, std::vector<unsigned int> > {};  // Thirdparty/DBoW2/DBoW2/FeatureVector.h
}
namespace ORB_SLAM {
class KeyFrame {
public:
    float fx, fy, cx, cy;
    cv::Mat GetRotation();
    cv::Mat GetTranslation();
    cv::Mat GetCameraCenter();
    DBoW2::FeatureVector GetFeatureVector();
    std::set<MapPoint *> GetMapPoints();
    std::vector<MapPoint *> GetMapPointMatches();
    MapPoint *GetMapPoint(const size_t &idx);
    void AddMapPoint(MapPoint *pMP, const size_t &idx);
    cv::KeyPoint GetKeyPointUn(const size_t &idx) const;
    cv::Mat GetDescriptor(const size_t &idx);
    int GetKeyPointScaleLevel(const size_t &idx) const;
    std::vector<cv::KeyPoint> GetKeyPointsUn() const;
    cv::Mat GetDescriptors();
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const;
    bool IsInImage(const float &x, const float &y) const;
    float GetScaleFactor(int nLevel = 1) const;
    std::vector<float> GetScaleFactors() const;
    float GetSigma2(int nLevel = 1) const;
    int GetScaleLevels() const;
    // ---- test-only state behind the getters above (the real class keeps the same data protected, KeyFrame.h:150-190) ----
    std::vector<cv::KeyPoint> t_keys;           // mvKeysUn
    cv::Mat t_desc;                             // mDescriptors
    std::vector<MapPoint *> t_mps;              // mvpMapPoints
    std::vector<float> t_sf;                    // mvScaleFactors
    cv::Mat t_Rcw, t_tcw, t_Ow;
    float t_minx, t_miny, t_maxx, t_maxy, t_ginvw, t_ginvh;   // mnMinX.., mfGridElementWidthInv..
    std::vector<size_t> t_grid[64][48];         // mGrid
    DBoW2::FeatureVector t_fv;
};
}
