// ORBmatcher.cc -- facade of ORB_SLAM::ORBmatcher (reference src/ORBmatcher.cc) over liborbfe.so.
//
// Pattern of every routine here: (1) gather -- walk the queries in the reference's order and enumerate their
// candidates (Frame's own GetFeaturesInArea, or the plain-array matchers of include/orbfe_match.h);
// (2) one GPU launch computes all 256-bit Hamming distances of the call; (3) replay -- the reference's
// sequential accept/skip loop runs over those distances, so results (incl. tie-breaks) are unchanged.
//
// Implemented: DescriptorDistance and the Frame-level routines of the Tracking thread; the array-level
// form of SearchByProjection(Frame,KeyFrame*,...) exists in the C-ABI (orbfe_search_by_projection_kf).
// Not yet implemented here (declared in the header, open work in DESIGN.md): the KeyFrame-typed routines
// SearchByProjection(Frame,KeyFrame*), SearchByProjection(KeyFrame*,Scw), SearchByBoW x2,
// SearchForTriangulation, SearchBySim3, Fuse x2.
#include "ORBmatcher.h"

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>

#include "orbfe.h"
#include "orbfe_match.h"

namespace ORB_SLAM {

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {

int g_match_device = 0;

// ORBmatcher objects are stack temporaries used from three threads (Tracking.cc:352,488,556,...):
// each calling thread lazily gets its own device handle (stream + scratch buffers).
OrbfeMatcher *thread_matcher() {
    static thread_local OrbfeMatcher *m = NULL;
    if (!m) {
        const int rc = orbfe_matcher_create(g_match_device, &m);
        if (rc != ORBFE_OK) {
            std::fprintf(stderr, "ORBmatcher: liborbfe error %d: %s (there is no CPU path)\n", rc, orbfe_last_error());
            std::abort();
        }
    }
    return m;
}

void check(int rc) {
    if (rc != ORBFE_OK) {
        std::fprintf(stderr, "ORBmatcher: liborbfe error %d: %s\n", rc, orbfe_last_error());
        std::abort();
    }
}

// contiguous copy of a frame's descriptors (cv::Mat rows may be strided)
struct DescBuf {
    std::vector<unsigned char> own;
    const unsigned char *ptr;
    explicit DescBuf(const cv::Mat &d) : ptr(NULL) {
        if (d.empty()) return;
        if (d.isContinuous()) { ptr = d.ptr(0); return; }
        own.resize((size_t)d.rows * 32);
        for (int i = 0; i < d.rows; i++) std::memcpy(&own[(size_t)i * 32], d.ptr(i), 32);
        ptr = &own[0];
    }
};

OrbfeFrameView make_view(const Frame &F, const DescBuf &d) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(OrbfeKeyPoint), "cv::KeyPoint must be the 28-byte OpenCV 2.4 layout");
    OrbfeFrameView v;
    v.n = (int)F.mvKeysUn.size();
    v.keys_un = v.n ? reinterpret_cast<const OrbfeKeyPoint *>(&F.mvKeysUn[0]) : NULL;
    v.desc = d.ptr;
    v.min_x = (float)Frame::mnMinX; v.min_y = (float)Frame::mnMinY;
    v.max_x = (float)Frame::mnMaxX; v.max_y = (float)Frame::mnMaxY;
    v.grid_inv_w = Frame::mfGridElementWidthInv;
    v.grid_inv_h = Frame::mfGridElementHeightInv;
    v.nlevels = F.mnScaleLevels;
    v.scale_factors = F.mvScaleFactors.empty() ? NULL : &F.mvScaleFactors[0];
    return v;
}

}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// A single pair is 8 XOR+popcount on the host: shipping 64 bytes to the GPU for one distance would only add
// latency.  Every *batched* distance computation below goes to the device.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    const unsigned char *pa = a.ptr(0), *pb = b.ptr(0);
    int dist = 0;
    for (int i = 0; i < 32; i += 8) {
        unsigned long long x, y;
        std::memcpy(&x, pa + i, 8);
        std::memcpy(&y, pb + i, 8);
        dist += __builtin_popcountll(x ^ y);
    }
    return dist;
}

float ORBmatcher::RadiusByViewingCos(const float &viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }  // :127-133

void ORBmatcher::ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// ---- Tracking::TrackWithMotionModel (Tracking.cc:565) -------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, float th)
{
    const DescBuf dc(CurrentFrame.mDescriptors), dl(LastFrame.mDescriptors);
    const OrbfeFrameView cur = make_view(CurrentFrame, dc), last = make_view(LastFrame, dl);
    const int nl = last.n, nc = cur.n;
    std::vector<unsigned char> has(nl, 0), outl(nl, 0);
    std::vector<float> world((size_t)nl * 3, 0.f);
    for (int i = 0; i < nl; i++) {
        MapPoint *pMP = LastFrame.mvpMapPoints[i];
        if (!pMP) continue;
        has[i] = 1;
        outl[i] = LastFrame.mvbOutlier[i] ? 1 : 0;
        const cv::Mat X = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) world[(size_t)i * 3 + k] = X.at<float>(k, 0);
    }
    float T[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
    std::vector<int> mp(nc, -1);
    for (int i = 0; i < nc; i++)
        if (CurrentFrame.mvpMapPoints[i]) mp[i] = INT_MAX;  // already occupied slot (:1562)
    const unsigned char *hp = nl ? &has[0] : NULL, *op = nl ? &outl[0] : NULL;
    const float *wp = nl ? &world[0] : NULL, *tp = T;
    int *mpp = nc ? &mp[0] : NULL;
    int nmatches = 0;
    check(orbfe_search_by_projection_frames(thread_matcher(), 1, &cur, &last, &hp, &op, &wp, &tp, Frame::fx, Frame::fy,
                                            Frame::cx, Frame::cy, th, mbCheckOrientation ? 1 : 0, &mpp, &nmatches));
    for (int i = 0; i < nc; i++)
        if (mp[i] >= 0 && mp[i] != INT_MAX) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[mp[i]];
    return nmatches;
}

// ---- Tracking::TrackPreviousFrame (Tracking.cc:497-502) ------------------------------------------------
int ORBmatcher::WindowSearch(Frame &F1, Frame &F2, int windowSize, std::vector<MapPoint *> &vpMapPointMatches2,
                             int minScaleLevel, int maxScaleLevel)
{
    const DescBuf d1(F1.mDescriptors), d2(F2.mDescriptors);
    const OrbfeFrameView v1 = make_view(F1, d1), v2 = make_view(F2, d2);
    std::vector<unsigned char> has(v1.n, 0);
    for (int i = 0; i < v1.n; i++) {
        MapPoint *p = F1.mvpMapPoints[i];
        has[i] = (p && !p->isBad()) ? 1 : 0;  // :425-428
    }
    std::vector<int> m21(v2.n > 0 ? v2.n : 1, -1);
    int nmatches = 0;
    check(orbfe_window_search(thread_matcher(), &v1, &v2, v1.n ? &has[0] : NULL, windowSize, minScaleLevel, maxScaleLevel,
                              mfNNratio, mbCheckOrientation ? 1 : 0, &m21[0], &nmatches));
    vpMapPointMatches2 = std::vector<MapPoint *>(F2.mvpMapPoints.size(), static_cast<MapPoint *>(NULL));  // :412
    for (int i2 = 0; i2 < v2.n; i2++)
        if (m21[i2] >= 0) vpMapPointMatches2[i2] = F1.mvpMapPoints[m21[i2]];
    return nmatches;
}

// ---- Tracking::Initialize (Tracking.cc:352-353) ----------------------------------------------------------
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched,
                                        std::vector<int> &vnMatches12, int windowSize)
{
    const DescBuf d1(F1.mDescriptors), d2(F2.mDescriptors);
    const OrbfeFrameView v1 = make_view(F1, d1), v2 = make_view(F2, d2);
    static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f must be two packed floats");
    vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);  // :601
    int nmatches = 0;
    if (v1.n == 0) return 0;
    check(orbfe_search_for_initialization(thread_matcher(), &v1, &v2, reinterpret_cast<float *>(&vbPrevMatched[0]), windowSize,
                                          mfNNratio, mbCheckOrientation ? 1 : 0, &vnMatches12[0], &nmatches));
    return nmatches;
}

// ---- Tracking::SearchReferencePointsInFrustum (Tracking.cc:724) -----------------------------------------
int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th)
{
    const DescBuf df(F.mDescriptors);
    const OrbfeFrameView fv = make_view(F, df);
    const int np = (int)vpMapPoints.size();
    std::vector<unsigned char> in_view(np, 0), desc((size_t)np * 32 + 1, 0);
    std::vector<float> proj((size_t)np * 2 + 1, 0.f), vcos(np + 1, 0.f);
    std::vector<int> level(np + 1, 0);
    for (int i = 0; i < np; i++) {
        MapPoint *pMP = vpMapPoints[i];
        if (!pMP->mbTrackInView || pMP->isBad()) continue;  // :57-61
        in_view[i] = 1;
        proj[2 * i] = pMP->mTrackProjX; proj[2 * i + 1] = pMP->mTrackProjY;
        level[i] = pMP->mnTrackScaleLevel;
        vcos[i] = pMP->mTrackViewCos;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(0), 32);
    }
    std::vector<int> mp(fv.n > 0 ? fv.n : 1, -1);
    for (int i = 0; i < fv.n; i++)
        if (F.mvpMapPoints[i]) mp[i] = INT_MAX;
    int nmatches = 0;
    check(orbfe_search_local_points(thread_matcher(), &fv, np, np ? &in_view[0] : NULL, &proj[0], &level[0], &vcos[0], &desc[0], th,
                                    mfNNratio, &mp[0], &nmatches));
    for (int i = 0; i < fv.n; i++)
        if (mp[i] >= 0 && mp[i] != INT_MAX) F.mvpMapPoints[i] = vpMapPoints[mp[i]];
    return nmatches;
}

// ---- Tracking::TrackPreviousFrame refinement (Tracking.cc:528-531) --------------------------------------
int ORBmatcher::SearchByProjection(Frame &F1, Frame &F2, int windowSize, std::vector<MapPoint *> &vpMapPointMatches2)
{
    vpMapPointMatches2 = F2.mvpMapPoints;  // :521
    const std::set<MapPoint *> found(vpMapPointMatches2.begin(), vpMapPointMatches2.end());
    const DescBuf d1(F1.mDescriptors), d2(F2.mDescriptors);
    const OrbfeFrameView v1 = make_view(F1, d1), v2 = make_view(F2, d2);
    std::vector<unsigned char> valid(v1.n + 1, 0);
    std::vector<float> world((size_t)v1.n * 3 + 1, 0.f);
    for (int i1 = 0; i1 < v1.n; i1++) {
        MapPoint *p = F1.mvpMapPoints[i1];
        if (!p || p->isBad() || found.count(p)) continue;  // :533-537
        valid[i1] = 1;
        const cv::Mat X = p->GetWorldPos();
        for (int k = 0; k < 3; k++) world[(size_t)i1 * 3 + k] = X.at<float>(k, 0);
    }
    float T[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[4 * r + c] = F2.mTcw.at<float>(r, c);
    std::vector<int> mp(v2.n > 0 ? v2.n : 1, -1);
    for (int i = 0; i < v2.n; i++)
        if (vpMapPointMatches2[i]) mp[i] = INT_MAX;
    int nmatches = 0;
    check(orbfe_search_by_projection_f1f2(thread_matcher(), &v1, &v2, &valid[0], &world[0], T, Frame::fx, Frame::fy, Frame::cx,
                                          Frame::cy, windowSize, mfNNratio, &mp[0], &nmatches));
    for (int i = 0; i < v2.n; i++)
        if (mp[i] >= 0 && mp[i] != INT_MAX) vpMapPointMatches2[i] = F1.mvpMapPoints[mp[i]];
    return nmatches;
}

}  // namespace ORB_SLAM
