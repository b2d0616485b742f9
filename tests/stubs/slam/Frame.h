// TEST STUB of the slice of ORB_SLAM::Frame (reference include/Frame.h:42-139) the front-end boundary touches.
#pragma once
#include <cstddef>
#include <vector>
#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "MapPoint.h"
#include "KeyFrame.h"
namespace ORB_SLAM {
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
class Frame {
public:
    Frame() : N(0), mnScaleLevels(0), mfScaleFactor(0) {}
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    cv::Mat mDescriptors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw;  // 4x4 CV_32F
    static float fx, fy, cx, cy;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    static int mnMinX, mnMaxX, mnMinY, mnMaxY;
    int mnScaleLevels;
    float mfScaleFactor;
    std::vector<float> mvScaleFactors;
    DBoW2::FeatureVector mFeatVec;
    // reference Frame.cc:200-265 (a restatement lives in the conformance TU; the real Frame.cc is linked unchanged)
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};
}
