// conformance.cc -- compile/link check of the drop-in boundary WITHOUT OpenCV/ROS: reproduces the call
// expressions the reference makes into ORBextractor / ORBmatcher (Frame.cc:60,92-93; Tracking.cc:111,126,
// 352-353,497,528,565,724; MapPoint.cc:224) against include/ORBextractor.h + include/ORBmatcher.h and the
// minimal cv:: / Frame / MapPoint stubs of tests/stubs.  With a CUDA device it also runs them end to end.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "orbfe.h"

using namespace ORB_SLAM;

float Frame::fx = 500.f, Frame::fy = 500.f, Frame::cx = 320.f, Frame::cy = 240.f;
float Frame::mfGridElementWidthInv = 0.f, Frame::mfGridElementHeightInv = 0.f;
int Frame::mnMinX = 0, Frame::mnMaxX = 640, Frame::mnMinY = 0, Frame::mnMaxY = 480;

// stand-in for the reference's own Frame::GetFeaturesInArea (Frame.cc:200-265), needed only because the stub Frame
// has no Frame.cc; grid filled by fill_frame() below like Frame.cc:116-123
std::vector<size_t> Frame::GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel, const int maxLevel) const {
    std::vector<size_t> out;
    int x0 = (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv); x0 = x0 < 0 ? 0 : x0;
    if (x0 >= FRAME_GRID_COLS) return out;
    int x1 = (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv); x1 = x1 > FRAME_GRID_COLS - 1 ? FRAME_GRID_COLS - 1 : x1;
    if (x1 < 0) return out;
    int y0 = (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv); y0 = y0 < 0 ? 0 : y0;
    if (y0 >= FRAME_GRID_ROWS) return out;
    int y1 = (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv); y1 = y1 > FRAME_GRID_ROWS - 1 ? FRAME_GRID_ROWS - 1 : y1;
    if (y1 < 0) return out;
    const bool check = !(minLevel == -1 && maxLevel == -1), same = check && minLevel == maxLevel;
    for (int ix = x0; ix <= x1; ix++)
        for (int iy = y0; iy <= y1; iy++)
            for (size_t j = 0; j < mGrid[ix][iy].size(); j++) {
                const cv::KeyPoint &kp = mvKeysUn[mGrid[ix][iy][j]];
                if (same) { if (kp.octave != minLevel) continue; }
                else if (check) { if (kp.octave < minLevel || kp.octave > maxLevel) continue; }
                if (std::fabs(kp.pt.x - x) > r || std::fabs(kp.pt.y - y) > r) continue;
                out.push_back(mGrid[ix][iy][j]);
            }
    return out;
}

// link-only definitions of the KeyFrame stub (the real KeyFrame.cc is linked in a real integration)
namespace ORB_SLAM {
cv::Mat KeyFrame::GetRotation() { return cv::Mat(); }
cv::Mat KeyFrame::GetTranslation() { return cv::Mat(); }
cv::Mat KeyFrame::GetCameraCenter() { return cv::Mat(); }
DBoW2::FeatureVector KeyFrame::GetFeatureVector() { return DBoW2::FeatureVector(); }
std::set<MapPoint *> KeyFrame::GetMapPoints() { return std::set<MapPoint *>(); }
std::vector<MapPoint *> KeyFrame::GetMapPointMatches() { return std::vector<MapPoint *>(); }
MapPoint *KeyFrame::GetMapPoint(const size_t &) { return NULL; }
void KeyFrame::AddMapPoint(MapPoint *, const size_t &) {}
cv::KeyPoint KeyFrame::GetKeyPointUn(const size_t &) const { return cv::KeyPoint(); }
cv::Mat KeyFrame::GetDescriptor(const size_t &) { return cv::Mat(); }
int KeyFrame::GetKeyPointScaleLevel(const size_t &) const { return 0; }
std::vector<cv::KeyPoint> KeyFrame::GetKeyPointsUn() const { return std::vector<cv::KeyPoint>(); }
cv::Mat KeyFrame::GetDescriptors() { return cv::Mat(); }
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float &, const float &, const float &) const { return std::vector<size_t>(); }
bool KeyFrame::IsInImage(const float &, const float &) const { return false; }
float KeyFrame::GetScaleFactor(int) const { return 1.f; }
std::vector<float> KeyFrame::GetScaleFactors() const { return std::vector<float>(1, 1.f); }
float KeyFrame::GetSigma2(int) const { return 1.f; }
int KeyFrame::GetScaleLevels() const { return 1; }
}

static void fill_frame(Frame &F, ORBextractor *ex, cv::Mat &im) {
    (*ex)(im, cv::Mat(), F.mvKeys, F.mDescriptors);  // Frame.cc:60
    F.N = (int)F.mvKeys.size();
    F.mvKeysUn = F.mvKeys;
    F.mvpMapPoints.assign(F.N, static_cast<MapPoint *>(NULL));
    F.mvbOutlier.assign(F.N, false);
    F.mnScaleLevels = ex->GetLevels();        // Frame.cc:92
    F.mfScaleFactor = ex->GetScaleFactor();   // Frame.cc:93
    F.mvScaleFactors.resize(F.mnScaleLevels);
    F.mvScaleFactors[0] = 1.0f;
    for (int i = 1; i < F.mnScaleLevels; i++) F.mvScaleFactors[i] = F.mvScaleFactors[i - 1] * F.mfScaleFactor;
    Frame::mfGridElementWidthInv = (float)FRAME_GRID_COLS / (float)(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = (float)FRAME_GRID_ROWS / (float)(Frame::mnMaxY - Frame::mnMinY);
    for (int i = 0; i < F.N; i++) {
        const int px = (int)std::floor((F.mvKeysUn[i].pt.x - Frame::mnMinX) * Frame::mfGridElementWidthInv + 0.5f);
        const int py = (int)std::floor((F.mvKeysUn[i].pt.y - Frame::mnMinY) * Frame::mfGridElementHeightInv + 0.5f);
        if (px >= 0 && px < FRAME_GRID_COLS && py >= 0 && py < FRAME_GRID_ROWS) F.mGrid[px][py].push_back(i);
    }
    F.mTcw.create(4, 4, CV_32F);
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) F.mTcw.at<float>(r, c) = r == c ? 1.f : 0.f;
}

int main() {
    // MapPoint.cc:224
    cv::Mat a(1, 32, CV_8U), b(1, 32, CV_8U);
    for (int i = 0; i < 32; i++) { a.at<unsigned char>(0, i) = (unsigned char)(i * 7); b.at<unsigned char>(0, i) = (unsigned char)(i * 7 ^ 0x0F); }
    if (ORBmatcher::DescriptorDistance(a, b) != 32 * 4 || ORBmatcher::DescriptorDistance(a, a) != 0) { std::puts("DescriptorDistance wrong"); return 1; }
    if (ORBmatcher::TH_HIGH != 100 || ORBmatcher::TH_LOW != 50 || ORBmatcher::HISTO_LENGTH != 30) return 1;
    if (orbfe_device_count() == 0) { std::puts("conformance: compile+link ok (no CUDA device: run skipped)"); return 0; }

    const int nFeatures = 1000, nLevels = 8, Score = 1, fastTh = 20;
    const float fScaleFactor = 1.2f;
    ORBextractor *mpORBextractor = new ORBextractor(nFeatures, fScaleFactor, nLevels, Score, fastTh);   // Tracking.cc:111
    ORBextractor *mpIniORBextractor = new ORBextractor(nFeatures * 2, 1.2, 8, Score, fastTh);           // Tracking.cc:126

    // deterministic textured image and a 3-px shifted copy
    cv::Mat im1(480, 640, CV_8UC1), im2(480, 640, CV_8UC1);
    unsigned s = 12345u;
    std::vector<int> coarse(81 * 61);
    for (size_t i = 0; i < coarse.size(); i++) { s = s * 1664525u + 1013904223u; coarse[i] = (s >> 24); }
    for (int y = 0; y < 480; y++)
        for (int x = 0; x < 640; x++) {
            s = s * 1664525u + 1013904223u;
            const int v = (coarse[(y / 8) * 81 + x / 8] * 3 + (int)((s >> 24) & 31)) / 3;
            im1.at<unsigned char>(y, x) = (unsigned char)(v > 255 ? 255 : v);
        }
    for (int y = 0; y < 480; y++)
        for (int x = 0; x < 640; x++) im2.at<unsigned char>(y, x) = im1.at<unsigned char>(y, (x + 637) % 640);  // shift by +3 px

    Frame mInitialFrame, mCurrentFrame;
    fill_frame(mInitialFrame, mpIniORBextractor, im1);
    fill_frame(mCurrentFrame, mpIniORBextractor, im2);
    std::printf("ini extractor: %d / %d keypoints\n", mInitialFrame.N, mCurrentFrame.N);
    if (mInitialFrame.N != 2000 || mCurrentFrame.N != 2000) return 2;

    std::vector<cv::Point2f> mvbPrevMatched(mInitialFrame.mvKeysUn.size());
    for (size_t i = 0; i < mvbPrevMatched.size(); i++) mvbPrevMatched[i] = mInitialFrame.mvKeysUn[i].pt;
    std::vector<int> mvIniMatches;
    ORBmatcher matcher(0.9, true);
    int nmatches = matcher.SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100);  // Tracking.cc:352-353
    std::printf("SearchForInitialization: %d matches\n", nmatches);
    if (nmatches < 50) return 3;

    // give every feature of the first frame a map point at depth 4 (camera 1 = world)
    Frame mLastFrame, F2;
    fill_frame(mLastFrame, mpORBextractor, im1);
    fill_frame(F2, mpORBextractor, im2);
    std::vector<MapPoint> points(mLastFrame.N);
    for (int i = 0; i < mLastFrame.N; i++) {
        MapPoint &p = points[i];
        p.mWorldPos.create(3, 1, CV_32F);
        p.mWorldPos.at<float>(0, 0) = (mLastFrame.mvKeysUn[i].pt.x - Frame::cx) / Frame::fx * 4.f;
        p.mWorldPos.at<float>(1, 0) = (mLastFrame.mvKeysUn[i].pt.y - Frame::cy) / Frame::fy * 4.f;
        p.mWorldPos.at<float>(2, 0) = 4.f;
        p.mDescriptor = mLastFrame.mDescriptors.row(i).clone();
        mLastFrame.mvpMapPoints[i] = &p;
    }
    F2.mTcw.at<float>(0, 3) = 3.f * 4.f / Frame::fx;  // +3 px shift
    std::vector<MapPoint *> vpMapPointMatches;
    ORBmatcher m2(0.9, true);
    int minOctave = -1;
    int n1 = m2.WindowSearch(mLastFrame, F2, 200, vpMapPointMatches, minOctave);           // Tracking.cc:497
    std::printf("WindowSearch: %d matches\n", n1);
    F2.mvpMapPoints = vpMapPointMatches;
    int n2 = m2.SearchByProjection(mLastFrame, F2, 15, vpMapPointMatches);                 // Tracking.cc:528
    std::printf("SearchByProjection(F1,F2,15): +%d matches\n", n2);
    Frame F3;
    fill_frame(F3, mpORBextractor, im2);
    F3.mTcw = F2.mTcw.clone();
    int n3 = m2.SearchByProjection(F3, mLastFrame, 15);                                    // Tracking.cc:565
    std::printf("SearchByProjection(Cur,Last,15): %d matches\n", n3);
    if (n1 < 100 || n3 < 300) return 4;
    // local-map points with cached projections (Frame::isInFrustum fills these; Tracking.cc:709-724)
    Frame F4;
    fill_frame(F4, mpORBextractor, im2);
    std::vector<MapPoint *> mvpLocalMapPoints;
    for (int i = 0; i < mLastFrame.N; i++) {
        MapPoint &p = points[i];
        p.mbTrackInView = true;
        p.mTrackProjX = mLastFrame.mvKeysUn[i].pt.x + 3.f;
        p.mTrackProjY = mLastFrame.mvKeysUn[i].pt.y;
        p.mnTrackScaleLevel = mLastFrame.mvKeysUn[i].octave;
        p.mTrackViewCos = 0.9999f;
        mvpLocalMapPoints.push_back(&p);
    }
    ORBmatcher m3(0.8);
    float th = 3;
    int n4 = m3.SearchByProjection(F4, mvpLocalMapPoints, th);                             // Tracking.cc:724
    std::printf("SearchByProjection(F,LocalMapPoints,3): %d matches\n", n4);
    if (n4 < 300) return 5;
    delete mpORBextractor;
    delete mpIniORBextractor;
    std::puts("conformance: run ok");
    return 0;
}
