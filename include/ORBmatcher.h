/*
 * ORBmatcher.h -- API-compatible replacement of ORB_SLAM's matcher header (raulmur/ORB_SLAM v1.0.1,
 * include/ORBmatcher.h:37-107).  Every public member keeps the reference's name, parameter types and
 * defaults so that Tracking.cc / LocalMapping.cc / LoopClosing.cc / MapPoint.cc compile against it unchanged.
 *
 * Implementation: orb_slam_b200/host/ORBmatcher.cc.  Candidate enumeration (Frame::GetFeaturesInArea order)
 * and each routine's sequential accept loop stay on the host; every batch of 256-bit Hamming distances is
 * computed by liborbfe.so on the GPU (include/orbfe_match.h).  All methods are defined; status per method in DESIGN.md.
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <climits>
#include <set>
#include <utility>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"

namespace ORB_SLAM {

class ORBmatcher {
public:
    static const int TH_LOW;        // 50
    static const int TH_HIGH;       // 100
    static const int HISTO_LENGTH;  // 30

    ORBmatcher(float nnratio = 0.6, bool checkOri = true);

    // Addition (not in the reference, whose matchers cannot fail): what happens when a liborbfe call fails.  The
    // default handler logs and the method returns 0 matches, map untouched; ORBFE_ABORT_ON_ERROR=1 makes it abort.
    typedef void (*ErrorHandler)(int code, const char *message);
    static void SetErrorHandler(ErrorHandler handler);  // NULL restores the default

    // popcount(a XOR b) over two 32-byte descriptor rows (reference ORBmatcher.cc:1794-1810)
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    // ---- Tracking thread -------------------------------------------------------------------------
    // local-map points already projected by Frame::isInFrustum -> features of F   (ORBmatcher.cc:49-125)
    int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th = 3);
    // map points of the previous frame projected with CurrentFrame.mTcw            (ORBmatcher.cc:1507-1620)
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, float th);
    // relocalisation refinement against a keyframe                                (ORBmatcher.cc:1622-1746)
    int SearchByProjection(Frame &CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*> &sAlreadyFound, float th, int ORBdist);
    // same-position window between two frames                                     (ORBmatcher.cc:409-516)
    int WindowSearch(Frame &F1, Frame &F2, int windowSize, std::vector<MapPoint *> &vpMapPointMatches2, int minOctave = -1, int maxOctave = INT_MAX);
    // window search around the projection of F1's points into F2                  (ORBmatcher.cc:519-594)
    int SearchByProjection(Frame &F1, Frame &F2, int windowSize, std::vector<MapPoint *> &vpMapPointMatches2);
    // two-view initialisation, level-0 features only                              (ORBmatcher.cc:598-713)
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize = 10);

    // ---- vocabulary-guided brute force ------------------------------------------------------------
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches);    // ORBmatcher.cc:155-284
    int SearchByBoW(KeyFrame *pKF1, KeyFrame* pKF2, std::vector<MapPoint*> &vpMatches12);   // ORBmatcher.cc:715-850

    // ---- LocalMapping / LoopClosing threads --------------------------------------------------------
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, std::vector<MapPoint*> &vpMatched, int th);  // :286-407
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<cv::KeyPoint> &vMatchedKeys1,
                               std::vector<cv::KeyPoint> &vMatchedKeys2, std::vector<std::pair<size_t, size_t> > &vMatchedPairs);  // :852-1014
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint *> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12, float th);  // :1267-1505
    int Fuse(KeyFrame* pKF, std::vector<MapPoint *> &vpMapPoints, float th = 2.5);                          // :1016-1134
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, float th = 2.5);           // :1136-1265

protected:
    bool CheckDistEpipolarLine(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const cv::Mat &F12, const KeyFrame *pKF);
    float RadiusByViewingCos(const float &viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3);

    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM

#endif  // ORBMATCHER_H
