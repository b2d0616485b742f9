// conformance.cc -- compile/link check of the drop-in boundary WITHOUT OpenCV/ROS: reproduces the call
// expressions the reference makes into ORBextractor / ORBmatcher (Frame.cc:60,92-93; Tracking.cc:111,126,
// 352-353,497,528,565,724; MapPoint.cc:224) against include/ORBextractor.h + include/ORBmatcher.h and the
// minimal cv:: / Frame / MapPoint stubs of tests/stubs.  With a CUDA device it also runs them end to end.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <set>
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

// functional stand-ins for the KeyFrame methods ORBmatcher calls (reference src/KeyFrame.cc; the real KeyFrame.cc is
// linked in a real integration).  GetFeaturesInArea / IsInImage follow KeyFrame.cc:612-657.
namespace ORB_SLAM {
cv::Mat KeyFrame::GetRotation() { return t_Rcw.clone(); }
cv::Mat KeyFrame::GetTranslation() { return t_tcw.clone(); }
cv::Mat KeyFrame::GetCameraCenter() { return t_Ow.clone(); }
DBoW2::FeatureVector KeyFrame::GetFeatureVector() { return t_fv; }
std::set<MapPoint *> KeyFrame::GetMapPoints() {
    std::set<MapPoint *> s;
    for (size_t i = 0; i < t_mps.size(); i++)
        if (t_mps[i] && !t_mps[i]->isBad()) s.insert(t_mps[i]);   // KeyFrame.cc:244-257
    return s;
}
std::vector<MapPoint *> KeyFrame::GetMapPointMatches() { return t_mps; }
MapPoint *KeyFrame::GetMapPoint(const size_t &idx) { return t_mps[idx]; }
void KeyFrame::AddMapPoint(MapPoint *pMP, const size_t &idx) { t_mps[idx] = pMP; }
cv::KeyPoint KeyFrame::GetKeyPointUn(const size_t &idx) const { return t_keys[idx]; }
cv::Mat KeyFrame::GetDescriptor(const size_t &idx) { return t_desc.row((int)idx).clone(); }
int KeyFrame::GetKeyPointScaleLevel(const size_t &idx) const { return t_keys[idx].octave; }
std::vector<cv::KeyPoint> KeyFrame::GetKeyPointsUn() const { return t_keys; }
cv::Mat KeyFrame::GetDescriptors() { return t_desc.clone(); }
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float &x, const float &y, const float &r) const {
    std::vector<size_t> out;
    int x0 = (int)std::floor((x - t_minx - r) * t_ginvw); x0 = x0 < 0 ? 0 : x0;
    if (x0 >= 64) return out;
    int x1 = (int)std::ceil((x - t_minx + r) * t_ginvw); x1 = x1 > 63 ? 63 : x1;
    if (x1 < 0) return out;
    int y0 = (int)std::floor((y - t_miny - r) * t_ginvh); y0 = y0 < 0 ? 0 : y0;
    if (y0 >= 48) return out;
    int y1 = (int)std::ceil((y - t_miny + r) * t_ginvh); y1 = y1 > 47 ? 47 : y1;
    if (y1 < 0) return out;
    for (int ix = x0; ix <= x1; ix++)
        for (int iy = y0; iy <= y1; iy++)
            for (size_t j = 0; j < t_grid[ix][iy].size(); j++) {
                const cv::KeyPoint &kp = t_keys[t_grid[ix][iy][j]];
                if (std::fabs(kp.pt.x - x) <= r && std::fabs(kp.pt.y - y) <= r) out.push_back(t_grid[ix][iy][j]);
            }
    return out;
}
bool KeyFrame::IsInImage(const float &x, const float &y) const { return x >= t_minx && x < t_maxx && y >= t_miny && y < t_maxy; }
float KeyFrame::GetScaleFactor(int nLevel) const { return t_sf[nLevel]; }
std::vector<float> KeyFrame::GetScaleFactors() const { return t_sf; }
float KeyFrame::GetSigma2(int nLevel) const { return t_sf[nLevel] * t_sf[nLevel]; }
int KeyFrame::GetScaleLevels() const { return (int)t_sf.size(); }
}

// ------------------------------------------------------------------------------------------------
// Plain-loop restatements (host popcount, no GPU, no candidate lists shared with the facade) of the KeyFrame-level
// routines, used to check the facade end to end: ORBmatcher.cc:286-405 (SearchByProjection with Scw), :1034-1130 (Fuse)
// and :1132-1234 (Fuse with Scw).  They read the reference line by line; the cv::Mat algebra is spelled out with
// OpenCV 2.4's semantics (gemm / dot / norm accumulate in double, Mat / scalar multiplies by the reciprocal).
// ------------------------------------------------------------------------------------------------
namespace ref {
static int ham(const cv::Mat &a, const cv::Mat &b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a.ptr(0)[i] ^ b.ptr(0)[i]));
    return d;
}
struct Cam { float R[9], t[3], Ow[3]; };
static Cam cam_from_Scw(const cv::Mat &Scw) {   // :297-301
    Cam c;
    double dd = 0;
    for (int k = 0; k < 3; k++) dd += (double)Scw.at<float>(0, k) * (double)Scw.at<float>(0, k);
    const float scw = (float)std::sqrt(dd);
    const float inv = (float)(1.0 / (double)scw);
    for (int r = 0; r < 3; r++) {
        for (int k = 0; k < 3; k++) c.R[3 * r + k] = Scw.at<float>(r, k) * inv;
        c.t[r] = Scw.at<float>(r, 3) * inv;
    }
    for (int k = 0; k < 3; k++)
        c.Ow[k] = (float)(((double)c.R[k] * c.t[0] + (double)c.R[3 + k] * c.t[1] + (double)c.R[6 + k] * c.t[2]) * -1.0);
    return c;
}
static Cam cam_from_kf(KeyFrame *kf) {
    Cam c;
    const cv::Mat R = kf->GetRotation(), t = kf->GetTranslation(), O = kf->GetCameraCenter();
    for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) c.R[3 * r + k] = R.at<float>(r, k); c.t[r] = t.at<float>(r, 0); c.Ow[r] = O.at<float>(r, 0); }
    return c;
}
// the projection gates shared by the three routines; returns the best keypoint (or -1) and its distance
static int best_for_point(KeyFrame *pKF, MapPoint *pMP, const Cam &c, float th, bool invz_double, const std::vector<MapPoint *> *skipMatched,
                          int &bestDist) {
    bestDist = INT_MAX;
    const cv::Mat Xw = pMP->GetWorldPos();
    const float X[3] = {Xw.at<float>(0, 0), Xw.at<float>(1, 0), Xw.at<float>(2, 0)};
    float Xc[3];
    for (int k = 0; k < 3; k++)
        Xc[k] = (float)(((double)c.R[3 * k] * X[0] + (double)c.R[3 * k + 1] * X[1] + (double)c.R[3 * k + 2] * X[2]) + (double)c.t[k]);
    if (Xc[2] < 0.0f) return -1;
    const float invz = invz_double ? (float)(1.0 / (double)Xc[2]) : 1.0f / Xc[2];
    const float x = Xc[0] * invz, y = Xc[1] * invz;
    const float u = pKF->fx * x + pKF->cx, v = pKF->fy * y + pKF->cy;
    if (!pKF->IsInImage(u, v)) return -1;
    const float maxD = pMP->GetMaxDistanceInvariance(), minD = pMP->GetMinDistanceInvariance();
    const float PO[3] = {X[0] - c.Ow[0], X[1] - c.Ow[1], X[2] - c.Ow[2]};
    const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist < minD || dist > maxD) return -1;
    const cv::Mat Pn = pMP->GetNormal();
    const double dot = (double)PO[0] * Pn.at<float>(0, 0) + (double)PO[1] * Pn.at<float>(1, 0) + (double)PO[2] * Pn.at<float>(2, 0);
    if (dot < 0.5 * dist) return -1;
    const float ratio = dist / minD;
    const std::vector<float> sf = pKF->GetScaleFactors();
    int lvl = 0;
    while (lvl < (int)sf.size() && sf[lvl] < ratio) lvl++;   // lower_bound
    const int nPredictedLevel = std::min(lvl, pKF->GetScaleLevels() - 1);
    const float radius = th * sf[nPredictedLevel];
    const std::vector<size_t> vIndices = pKF->GetFeaturesInArea(u, v, radius);
    int bestIdx = -1;
    const cv::Mat dMP = pMP->GetDescriptor();
    for (size_t k = 0; k < vIndices.size(); k++) {
        const size_t idx = vIndices[k];
        if (skipMatched && (*skipMatched)[idx]) continue;
        const int kpLevel = pKF->GetKeyPointScaleLevel(idx);
        if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
        const int d = ham(dMP, pKF->GetDescriptor(idx));
        if (d < bestDist) { bestDist = d; bestIdx = (int)idx; }
    }
    return bestIdx;
}
static int SearchByProjection(KeyFrame *pKF, const cv::Mat &Scw, const std::vector<MapPoint *> &vpPoints, std::vector<MapPoint *> &vpMatched, int th) {
    const Cam c = cam_from_Scw(Scw);
    std::set<MapPoint *> found(vpMatched.begin(), vpMatched.end());
    found.erase(static_cast<MapPoint *>(NULL));
    int nmatches = 0;
    for (size_t i = 0; i < vpPoints.size(); i++) {
        MapPoint *pMP = vpPoints[i];
        if (pMP->isBad() || found.count(pMP)) continue;
        int bd;
        const int bi = best_for_point(pKF, pMP, c, (float)th, false, &vpMatched, bd);
        if (bd <= ORBmatcher::TH_LOW) { vpMatched[bi] = pMP; nmatches++; }
    }
    return nmatches;
}
static int Fuse(KeyFrame *pKF, std::vector<MapPoint *> &vpMapPoints, float th) {
    const Cam c = cam_from_kf(pKF);
    int nFused = 0;
    for (size_t i = 0; i < vpMapPoints.size(); i++) {
        MapPoint *pMP = vpMapPoints[i];
        if (!pMP) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        int bd;
        const int bi = best_for_point(pKF, pMP, c, th, false, NULL, bd);
        if (bd <= ORBmatcher::TH_LOW) {
            MapPoint *in = pKF->GetMapPoint(bi);
            if (in) { if (!in->isBad()) pMP->Replace(in); }
            else { pMP->AddObservation(pKF, bi); pKF->AddMapPoint(pMP, bi); }
            nFused++;
        }
    }
    return nFused;
}
static int FuseScw(KeyFrame *pKF, const cv::Mat &Scw, const std::vector<MapPoint *> &vpPoints, float th) {
    const Cam c = cam_from_Scw(Scw);
    const std::set<MapPoint *> found = pKF->GetMapPoints();
    int nFused = 0;
    for (size_t i = 0; i < vpPoints.size(); i++) {
        MapPoint *pMP = vpPoints[i];
        if (pMP->isBad() || found.count(pMP)) continue;
        int bd;
        const int bi = best_for_point(pKF, pMP, c, th, true, NULL, bd);
        if (bd <= ORBmatcher::TH_LOW) {
            MapPoint *in = pKF->GetMapPoint(bi);
            if (in) { if (!in->isBad()) in->Replace(pMP); }
            else { pMP->AddObservation(pKF, bi); pKF->AddMapPoint(pMP, bi); }
            nFused++;
        }
    }
    return nFused;
}
// SearchBySim3, ORBmatcher.cc:1264-1451: both projection directions, then the mutual-agreement filter
static void sim3_direction(KeyFrame *from, KeyFrame *to, const float Rw[9], const float tw[3], const float sR[9], const float t[3],
                           const std::vector<MapPoint *> &vp, const std::vector<bool> &done, float fx, float fy, float cx, float cy,
                           float th, std::vector<int> &match) {
    (void)from;
    const std::vector<float> sf = to->GetScaleFactors();
    const int nMaxLevel = to->GetScaleLevels() - 1;
    for (size_t i = 0; i < vp.size(); i++) {
        MapPoint *pMP = vp[i];
        if (!pMP || done[i]) continue;
        if (pMP->isBad()) continue;
        const cv::Mat Xw = pMP->GetWorldPos();
        float a[3], b[3];
        for (int k = 0; k < 3; k++)
            a[k] = (float)(((double)Rw[3 * k] * Xw.at<float>(0, 0) + (double)Rw[3 * k + 1] * Xw.at<float>(1, 0) + (double)Rw[3 * k + 2] * Xw.at<float>(2, 0)) + (double)tw[k]);
        for (int k = 0; k < 3; k++)
            b[k] = (float)(((double)sR[3 * k] * a[0] + (double)sR[3 * k + 1] * a[1] + (double)sR[3 * k + 2] * a[2]) + (double)t[k]);
        if (b[2] < 0.0) continue;
        const float invz = (float)(1.0 / b[2]);
        const float x = b[0] * invz, y = b[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!to->IsInImage(u, v)) continue;
        const float maxD = pMP->GetMaxDistanceInvariance(), minD = pMP->GetMinDistanceInvariance();
        const float dist3D = (float)std::sqrt((double)b[0] * b[0] + (double)b[1] * b[1] + (double)b[2] * b[2]);
        if (dist3D < minD || dist3D > maxD) continue;
        const float ratio = dist3D / minD;
        int lvl = 0;
        while (lvl < (int)sf.size() && sf[lvl] < ratio) lvl++;
        const int nPredictedLevel = std::min(lvl, nMaxLevel);
        const float radius = th * sf[nPredictedLevel];
        const std::vector<size_t> vIndices = to->GetFeaturesInArea(u, v, radius);
        if (vIndices.empty()) continue;
        const cv::Mat dMP = pMP->GetDescriptor();
        int bestDist = INT_MAX, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const cv::KeyPoint kp = to->GetKeyPointUn(vIndices[k]);
            if (kp.octave < nPredictedLevel - 1 || kp.octave > nPredictedLevel) continue;
            const int d = ham(dMP, to->GetDescriptor(vIndices[k]));
            if (d < bestDist) { bestDist = d; bestIdx = (int)vIndices[k]; }
        }
        if (bestDist <= ORBmatcher::TH_HIGH) match[i] = bestIdx;
    }
}
static int SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12, float s12, const cv::Mat &R12,
                        const cv::Mat &t12, float th) {
    const Cam c1 = cam_from_kf(pKF1), c2 = cam_from_kf(pKF2);
    float sR12[9], sR21[9], t21[3], t12f[3];
    const float inv_s = (float)(1.0 / (double)s12);   // (1.0/s12) * Mat: the double scalar multiplies in double, one rounding
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) {
            sR12[3 * r + k] = (float)((double)R12.at<float>(r, k) * (double)s12);
            sR21[3 * r + k] = R12.at<float>(k, r) * inv_s;
        }
    for (int k = 0; k < 3; k++) t12f[k] = t12.at<float>(k, 0);
    for (int k = 0; k < 3; k++)
        t21[k] = (float)(((double)sR21[3 * k] * t12f[0] + (double)sR21[3 * k + 1] * t12f[1] + (double)sR21[3 * k + 2] * t12f[2]) * -1.0);
    const std::vector<MapPoint *> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vp1.size(), N2 = (int)vp2.size();
    std::vector<bool> done1(N1, false), done2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint *pMP = vpMatches12[i];
        if (pMP) {
            done1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) done2[idx2] = true;
        }
    }
    std::vector<int> m1(N1, -1), m2(N2, -1);
    sim3_direction(pKF1, pKF2, c1.R, c1.t, sR21, t21, vp1, done1, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, th, m1);
    sim3_direction(pKF2, pKF1, c2.R, c2.t, sR12, t12f, vp2, done2, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, th, m2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = m1[i1];
        if (idx2 >= 0) {
            const int idx1 = m2[idx2];
            if (idx1 == i1) { vpMatches12[i1] = vp2[idx2]; nFound++; }
        }
    }
    return nFound;
}
}  // namespace ref

// a keyframe from a frame's features, observed from pose [R|t] = [I | (tx,0,0)]; `occupied`: every n-th slot already
// holds a (foreign) map point
static void make_keyframe(KeyFrame &kf, const Frame &F, float tx, std::vector<MapPoint> &foreign, int every) {
    kf.fx = Frame::fx; kf.fy = Frame::fy; kf.cx = Frame::cx; kf.cy = Frame::cy;
    kf.t_keys = F.mvKeysUn;
    kf.t_desc = F.mDescriptors.clone();
    kf.t_sf = F.mvScaleFactors;
    kf.t_minx = (float)Frame::mnMinX; kf.t_miny = (float)Frame::mnMinY; kf.t_maxx = (float)Frame::mnMaxX; kf.t_maxy = (float)Frame::mnMaxY;
    kf.t_ginvw = Frame::mfGridElementWidthInv; kf.t_ginvh = Frame::mfGridElementHeightInv;
    for (int ix = 0; ix < 64; ix++) for (int iy = 0; iy < 48; iy++) kf.t_grid[ix][iy] = F.mGrid[ix][iy];
    kf.t_Rcw.create(3, 3, CV_32F); kf.t_tcw.create(3, 1, CV_32F); kf.t_Ow.create(3, 1, CV_32F);
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) kf.t_Rcw.at<float>(r, c) = r == c ? 1.f : 0.f; kf.t_tcw.at<float>(r, 0) = 0.f; kf.t_Ow.at<float>(r, 0) = 0.f; }
    kf.t_tcw.at<float>(0, 0) = tx; kf.t_Ow.at<float>(0, 0) = -tx;
    kf.t_mps.assign(F.N, static_cast<MapPoint *>(NULL));
    foreign.assign(F.N / every + 1, MapPoint());
    for (int i = 0, k = 0; i < F.N; i += every, k++) kf.t_mps[i] = &foreign[k];
}

// KeyFrame-level facade methods against the plain-loop restatements above, on identical copies of the scene
static int check_keyframe_routines(const Frame &F1, const Frame &F2) {
    const float tx = 3.f * 4.f / Frame::fx;  // F2 sees the scene shifted by +3 px
    // map points from F1's features at depth 4 (camera 1 = world), ORB-SLAM's UpdateNormalAndDepth conventions
    struct Scene {
        std::vector<MapPoint> pts, foreign;
        KeyFrame kf;
        std::vector<MapPoint *> vp;
    };
    Scene A, B;  // A: facade, B: restatement
    Scene *S[2] = {&A, &B};
    for (int s = 0; s < 2; s++) {
        Scene &sc = *S[s];
        make_keyframe(sc.kf, F2, tx, sc.foreign, 9);
        sc.pts.assign(F1.N, MapPoint());
        for (int i = 0; i < F1.N; i++) {
            MapPoint &p = sc.pts[i];
            p.mWorldPos.create(3, 1, CV_32F);
            const float X = (F1.mvKeysUn[i].pt.x - Frame::cx) / Frame::fx * 4.f, Y = (F1.mvKeysUn[i].pt.y - Frame::cy) / Frame::fy * 4.f, Z = 4.f;
            p.mWorldPos.at<float>(0, 0) = X; p.mWorldPos.at<float>(1, 0) = Y; p.mWorldPos.at<float>(2, 0) = Z;
            const float d = std::sqrt(X * X + Y * Y + Z * Z);
            p.mNormal.create(3, 1, CV_32F);
            p.mNormal.at<float>(0, 0) = X / d; p.mNormal.at<float>(1, 0) = Y / d; p.mNormal.at<float>(2, 0) = Z / d;
            const int lv = F1.mvKeysUn[i].octave;
            // scale-invariance range in the spirit of MapPoint.cc:300-310, chosen so that the level predicted from
            // dist / minDistance is lv + 1 (accepting key points of level lv and lv + 1) for most points, lv + 2 for some
            p.mfMinDistance = d / (F1.mvScaleFactors[lv] * (i % 7 == 0 ? 1.25f : 1.05f));
            p.mfMaxDistance = (i % 11 == 0 ? 0.999f : 1.2f) * d * F1.mvScaleFactors[F1.mnScaleLevels - 1 - lv];
            p.mDescriptor = F1.mDescriptors.row(i).clone();
            p.mbBad = (i % 53 == 0);
            sc.vp.push_back(&p);
        }
    }
    ORBmatcher matcher(0.75, true);
    int rc = 0;
    // ---- SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), ORBmatcher.cc:286-405 ----
    {
        cv::Mat Scw(4, 4, CV_32F);
        const float s = 1.07f;   // a similarity with scale: sR = s*I, st = s*t
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Scw.at<float>(r, c) = r == c ? (r < 3 ? s : 1.f) : 0.f;
        Scw.at<float>(0, 3) = s * tx;
        std::vector<MapPoint *> mA(A.kf.t_mps), mB(B.kf.t_mps);
        const int nA = matcher.SearchByProjection(&A.kf, Scw, A.vp, mA, 10);
        const int nB = ref::SearchByProjection(&B.kf, Scw, B.vp, mB, 10);
        int diff = 0, newm = 0;
        for (size_t i = 0; i < mA.size(); i++) {
            const long ia = mA[i] && mA[i] >= &A.pts[0] && mA[i] < &A.pts[0] + A.pts.size() ? (long)(mA[i] - &A.pts[0]) : (mA[i] ? -2 : -1);
            const long ib = mB[i] && mB[i] >= &B.pts[0] && mB[i] < &B.pts[0] + B.pts.size() ? (long)(mB[i] - &B.pts[0]) : (mB[i] ? -2 : -1);
            if (ia != ib) diff++;
            if (ia >= 0) newm++;
        }
        std::printf("SearchByProjection(KF,Scw): facade %d, restatement %d, %d new, %d slots differ\n", nA, nB, newm, diff);
        if (nA != nB || diff || nA < 200) rc = 10;
    }
    // ---- Fuse(KeyFrame*, vpMapPoints, th), ORBmatcher.cc:1034-1130 ----
    {
        const int nA = matcher.Fuse(&A.kf, A.vp, 3.0f);
        const int nB = ref::Fuse(&B.kf, B.vp, 3.0f);
        int diff = 0, repl = 0, added = 0;
        for (size_t i = 0; i < A.pts.size(); i++) {
            const bool ra = A.pts[i].t_replaced_by != NULL, rb = B.pts[i].t_replaced_by != NULL;
            if (ra != rb) diff++;
            else if (ra && (A.pts[i].t_replaced_by - &A.foreign[0]) != (B.pts[i].t_replaced_by - &B.foreign[0])) diff++;
            if (A.pts[i].GetIndexInKeyFrame(&A.kf) != B.pts[i].GetIndexInKeyFrame(&B.kf)) diff++;
            repl += ra; added += A.pts[i].IsInKeyFrame(&A.kf);
        }
        std::printf("Fuse(KF,points): facade %d, restatement %d (%d replaced, %d added), %d differences\n", nA, nB, repl, added, diff);
        if (nA != nB || diff || nA < 200 || repl == 0 || added == 0) rc = rc ? rc : 11;
    }
    // ---- Fuse(KeyFrame*, Scw, vpPoints, th), ORBmatcher.cc:1132-1234, on fresh copies ----
    {
        Scene C, D;
        Scene *T[2] = {&C, &D};
        for (int s = 0; s < 2; s++) {
            make_keyframe(T[s]->kf, F2, tx, T[s]->foreign, 5);
            T[s]->pts = A.pts;   // same geometry; reset the bookkeeping
            for (size_t i = 0; i < T[s]->pts.size(); i++) { T[s]->pts[i].t_obs.clear(); T[s]->pts[i].t_replaced_by = NULL; T[s]->pts[i].mbBad = (i % 41 == 0); T[s]->vp.push_back(&T[s]->pts[i]); }
        }
        cv::Mat Scw(4, 4, CV_32F);
        const float s = 0.93f;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Scw.at<float>(r, c) = r == c ? (r < 3 ? s : 1.f) : 0.f;
        Scw.at<float>(0, 3) = s * tx;
        const int nC = matcher.Fuse(&C.kf, Scw, C.vp, 4.0f);
        const int nD = ref::FuseScw(&D.kf, Scw, D.vp, 4.0f);
        int diff = 0, repl = 0;
        for (size_t i = 0; i < C.foreign.size(); i++) {
            const bool rc_ = C.foreign[i].t_replaced_by != NULL, rd = D.foreign[i].t_replaced_by != NULL;
            if (rc_ != rd) diff++;
            else if (rc_ && (C.foreign[i].t_replaced_by - &C.pts[0]) != (D.foreign[i].t_replaced_by - &D.pts[0])) diff++;
            repl += rc_;
        }
        for (size_t i = 0; i < C.pts.size(); i++)
            if (C.pts[i].GetIndexInKeyFrame(&C.kf) != D.pts[i].GetIndexInKeyFrame(&D.kf)) diff++;
        std::printf("Fuse(KF,Scw,points): facade %d, restatement %d (%d replaced), %d differences\n", nC, nD, repl, diff);
        if (nC != nD || diff || nC < 200 || repl == 0) rc = rc ? rc : 12;
    }
    // ---- SearchBySim3(KF1, KF2, vpMatches12, s12, R12, t12, th), ORBmatcher.cc:1264-1451 (no mutations: one scene) ----
    {
        KeyFrame kf1, kf2;
        std::vector<MapPoint> unused1, unused2;
        make_keyframe(kf1, F1, 0.f, unused1, 1 << 30);
        make_keyframe(kf2, F2, tx, unused2, 1 << 30);
        const Frame *Fs[2] = {&F1, &F2};
        KeyFrame *kfs[2] = {&kf1, &kf2};
        const float txs[2] = {0.f, tx};
        std::vector<MapPoint> pts[2];
        for (int s = 0; s < 2; s++) {
            const Frame &F = *Fs[s];
            pts[s].assign(F.N, MapPoint());
            for (int i = 0; i < F.N; i++) {
                MapPoint &p = pts[s][i];
                const float Xc = (F.mvKeysUn[i].pt.x - Frame::cx) / Frame::fx * 4.f, Yc = (F.mvKeysUn[i].pt.y - Frame::cy) / Frame::fy * 4.f;
                p.mWorldPos.create(3, 1, CV_32F);
                p.mWorldPos.at<float>(0, 0) = Xc - txs[s]; p.mWorldPos.at<float>(1, 0) = Yc; p.mWorldPos.at<float>(2, 0) = 4.f;
                const float d = std::sqrt(Xc * Xc + Yc * Yc + 16.f);
                const int lv = F.mvKeysUn[i].octave;
                p.mfMinDistance = d / (F.mvScaleFactors[lv] * (i % 5 == 0 ? 1.3f : 1.05f));
                p.mfMaxDistance = (i % 13 == 0 ? 0.999f : 1.2f) * d * F.mvScaleFactors[F.mnScaleLevels - 1 - lv];
                p.mDescriptor = F.mDescriptors.row(i).clone();
                p.mbBad = (i % 37 == 0);
                p.t_obs[kfs[s]] = (size_t)i;
                if (i % 4 != 3) kfs[s]->t_mps[i] = &p;   // a quarter of the features have no map point
            }
        }
        cv::Mat R12(3, 3, CV_32F), t12(3, 1, CV_32F);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R12.at<float>(r, c) = r == c ? 1.f : 0.f;
        const float s12 = 1.01f;
        t12.at<float>(0, 0) = -tx; t12.at<float>(1, 0) = 0.f; t12.at<float>(2, 0) = 0.02f;
        std::vector<MapPoint *> mA(F1.N, static_cast<MapPoint *>(NULL));
        for (int i = 0; i < F1.N && i < F2.N; i += 17) mA[i] = &pts[1][(i * 7) % F2.N];   // already matched pairs
        std::vector<MapPoint *> mB(mA);
        const int nA = matcher.SearchBySim3(&kf1, &kf2, mA, s12, R12, t12, 7.5f);
        const int nB = ref::SearchBySim3(&kf1, &kf2, mB, s12, R12, t12, 7.5f);
        int diff = 0;
        for (size_t i = 0; i < mA.size(); i++) diff += mA[i] != mB[i];
        std::printf("SearchBySim3: facade %d, restatement %d, %d slots differ\n", nA, nB, diff);
        if (nA != nB || diff || nA < 150) rc = rc ? rc : 13;
    }
    return rc;
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
    if (orbfe_device_count() == 0) {
        // No device: the error policy of the facades is what can be checked here.  The reference's constructor and
        // operator() cannot fail; ours report through the handler and return no keypoints / 0 matches instead of
        // aborting the process (ORBFE_ABORT_ON_ERROR=1 restores the abort).
        static int n_err = 0;
        struct H { static void on_error(int, const char *) { n_err++; } };
        ORBextractor::SetErrorHandler(&H::on_error);
        ORBmatcher::SetErrorHandler(&H::on_error);
        ORBextractor ex(500, 1.2f, 8, 1, 20);             // create fails: no device
        cv::Mat im(120, 160, CV_8UC1);
        for (int y = 0; y < 120; y++) for (int x = 0; x < 160; x++) im.at<unsigned char>(y, x) = (unsigned char)((x * 7) ^ (y * 13));
        std::vector<cv::KeyPoint> kp(3);
        cv::Mat desc;
        ex(im, cv::Mat(), kp, desc);                      // reports again, returns nothing
        if (n_err < 2 || !kp.empty() || !desc.empty()) { std::printf("error policy: n_err=%d kp=%d\n", n_err, (int)kp.size()); return 9; }
        ORBextractor::SetErrorHandler(NULL);
        ORBmatcher::SetErrorHandler(NULL);
        std::puts("conformance: compile+link ok, error policy ok (no CUDA device: run skipped)");
        return 0;
    }

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
    // KeyFrame-level routines (loop closing / local mapping): facade vs plain-loop restatement on functional stubs
    {
        Frame K1, K2;
        fill_frame(K1, mpORBextractor, im1);
        fill_frame(K2, mpORBextractor, im2);
        const int krc = check_keyframe_routines(K1, K2);
        if (krc) return krc;
    }
    delete mpORBextractor;
    delete mpIniORBextractor;
    std::puts("conformance: run ok");
    return 0;
}
