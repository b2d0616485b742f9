// ORBmatcher.cc -- facade of ORB_SLAM::ORBmatcher (reference src/ORBmatcher.cc) over liborbfe.so.
//
// Pattern of every routine here: (1) gather -- walk the queries in the reference's order and enumerate their
// candidates (Frame's own GetFeaturesInArea, or the plain-array matchers of include/orbfe_match.h);
// (2) one GPU launch computes all 256-bit Hamming distances of the call; (3) replay -- the reference's
// sequential accept/skip loop runs over those distances, so results (incl. tie-breaks) are unchanged.
//
// Implemented: DescriptorDistance and the Frame-level routines of the Tracking thread.
// Not yet implemented (declared in the header, open work in DESIGN.md): the KeyFrame-level routines
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
// queries = local map points with their cached projection; candidates through Frame::GetFeaturesInArea.
int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th)
{
    const bool bFactor = th != 1.0;
    const DescBuf df(F.mDescriptors);
    std::vector<int> row_ptr(1, 0), cols, qmp;
    std::vector<unsigned char> qdesc;
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint *pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView) continue;
        if (pMP->isBad()) continue;
        const int nPredictedLevel = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);
        if (bFactor) r *= th;
        const std::vector<size_t> vNear = F.GetFeaturesInArea(pMP->mTrackProjX, pMP->mTrackProjY,
                                                              r * F.mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel);
        if (vNear.empty()) continue;
        const cv::Mat d = pMP->GetDescriptor();
        qdesc.insert(qdesc.end(), d.ptr(0), d.ptr(0) + 32);
        for (size_t k = 0; k < vNear.size(); k++) cols.push_back((int)vNear[k]);
        row_ptr.push_back((int)cols.size());
        qmp.push_back((int)iMP);
    }
    std::vector<unsigned short> dist(cols.size() ? cols.size() : 1);
    if (!cols.empty())
        check(orbfe_hamming_csr(thread_matcher(), &qdesc[0], (int)qmp.size(), df.ptr, F.mDescriptors.rows, &row_ptr[0], &cols[0], &dist[0]));
    int nmatches = 0;
    for (size_t q = 0; q < qmp.size(); q++) {  // :82-121
        int bestDist = INT_MAX, bestLevel = -1, bestDist2 = INT_MAX, bestLevel2 = -1, bestIdx = -1;
        for (int c = row_ptr[q]; c < row_ptr[q + 1]; c++) {
            const int idx = cols[c];
            if (F.mvpMapPoints[idx]) continue;
            const int d = dist[c];
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = F.mvKeysUn[idx].octave; bestIdx = idx; }
            else if (d < bestDist2) { bestLevel2 = F.mvKeysUn[idx].octave; bestDist2 = d; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
            F.mvpMapPoints[bestIdx] = vpMapPoints[qmp[q]];
            nmatches++;
        }
    }
    return nmatches;
}

// ---- Tracking::TrackPreviousFrame refinement (Tracking.cc:528-531) --------------------------------------
int ORBmatcher::SearchByProjection(Frame &F1, Frame &F2, int windowSize, std::vector<MapPoint *> &vpMapPointMatches2)
{
    vpMapPointMatches2 = F2.mvpMapPoints;  // :521
    const std::set<MapPoint *> found(vpMapPointMatches2.begin(), vpMapPointMatches2.end());
    const DescBuf d1(F1.mDescriptors), d2(F2.mDescriptors);
    std::vector<int> row_ptr(1, 0), cols, q1;
    for (size_t i1 = 0; i1 < F1.mvpMapPoints.size(); i1++) {
        MapPoint *pMP1 = F1.mvpMapPoints[i1];
        if (!pMP1) continue;
        if (pMP1->isBad() || found.count(pMP1)) continue;
        const int level1 = F1.mvKeysUn[i1].octave;
        // x3Dc2 = Rc2w*x3Dw + tc2w: cv::gemm on CV_32F accumulates and adds in double (:540-541)
        const cv::Mat X = pMP1->GetWorldPos();
        float xc[3];
        for (int k = 0; k < 3; k++) {
            const double s = (double)F2.mTcw.at<float>(k, 0) * (double)X.at<float>(0, 0) + (double)F2.mTcw.at<float>(k, 1) * (double)X.at<float>(1, 0) +
                             (double)F2.mTcw.at<float>(k, 2) * (double)X.at<float>(2, 0);
            xc[k] = (float)(s + (double)F2.mTcw.at<float>(k, 3));
        }
        const float invz = 1.0 / xc[2];
        const float u2 = Frame::fx * xc[0] * invz + Frame::cx;
        const float v2 = Frame::fy * xc[1] * invz + Frame::cy;
        const std::vector<size_t> vIdx = F2.GetFeaturesInArea(u2, v2, windowSize, level1, level1);
        if (vIdx.empty()) continue;
        for (size_t k = 0; k < vIdx.size(); k++) cols.push_back((int)vIdx[k]);
        row_ptr.push_back((int)cols.size());
        q1.push_back((int)i1);
    }
    // distances: queries are rows of F1's descriptor matrix -> gather them contiguously
    std::vector<unsigned char> qd(q1.size() * 32 + 1);
    for (size_t q = 0; q < q1.size(); q++) std::memcpy(&qd[q * 32], d1.ptr + (size_t)q1[q] * 32, 32);
    std::vector<unsigned short> dist(cols.size() ? cols.size() : 1);
    if (!cols.empty())
        check(orbfe_hamming_csr(thread_matcher(), &qd[0], (int)q1.size(), d2.ptr, F2.mDescriptors.rows, &row_ptr[0], &cols[0], &dist[0]));
    int nmatches = 0;
    for (size_t q = 0; q < q1.size(); q++) {  // :561-590
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = row_ptr[q]; c < row_ptr[q + 1]; c++) {
            const int i2 = cols[c];
            if (vpMapPointMatches2[i2]) continue;
            const int d = dist[c];
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (static_cast<float>(bestDist) <= static_cast<float>(bestDist2) * mfNNratio && bestDist <= TH_HIGH) {
            vpMapPointMatches2[bestIdx2] = F1.mvpMapPoints[q1[q]];
            nmatches++;
        }
    }
    return nmatches;
}

}  // namespace ORB_SLAM
