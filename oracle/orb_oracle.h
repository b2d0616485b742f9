/*
 * orb_oracle.h -- CPU ORACLE for the ORB feature front-end.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's CPU algorithm (raulmur/ORB_SLAM v1.0.1,
 * src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc) plus the OpenCV primitives those files call
 * (OpenCV is NOT vendored in the reference; semantics restated per SURVEY.md section 8c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * The product (orb_slam_b200/, liborbfe.so) never links, imports or calls anything in oracle/.
 *
 * PARITY PIN STATUS: the reference has no tests / golden vectors for this path and cannot be compiled
 * here (needs OpenCV 2.4 + ROS + Boost).  The OpenCV primitives restated here (resize, FAST, integer
 * GaussianBlur engine, fastAtan2, copyMakeBorder, undistortPoints) are pinned bit-for-bit against python cv2 4.13 by
 * tests/golden/ (generator: tests/golden/make_golden.py; tests/test_oracle_golden.py).
 * The extractor's pipeline glue (cell grid, quota, retention, orientation, descriptor) is pinned end to end against a
 * Python composition of LIVE cv2 calls that follows ComputeKeyPoints / operator() line by line
 * (test_extract_composition_against_live_cv2): identical keypoints, responses, order, angles and descriptors.
 * The matcher / BoW / keyframe-database routines have no executable reference and no golden vectors ("parity
 * unpinned", stated as such in DESIGN.md); each is cross-checked against an independent pure-Python loop written
 * from the reference source (tests/test_oracle_match.py, tests/test_oracle_bow.py).
 *
 * Canonical choices where the reference itself is toolchain-dependent (see DESIGN.md "Canonical semantics"):
 *   - GaussianBlur: OpenCV-2.4 integer engine, taps [18,34,49,55,49,34,18]/256 per pass, /65536 half-even.
 *   - retainBest ties: top-n by response, ties at the cut broken by earlier position
 *     (ORB_ORACLE_TIES_CANONICAL); ORB_ORACLE_TIES_NTH_ELEMENT calls std::nth_element literally.
 *   - float arithmetic: binary32, every op rounded, no FMA contraction (-ffp-contract=off).
 *   - cos/sin of the keypoint angle: correctly-rounded float, (float)cos((double)a)
 *     (ORB_ORACLE_TRIG_RN); ORB_ORACLE_TRIG_LIBMF uses cosf/sinf which is what GCC's overload
 *     resolution picks in the reference TU (differs from RN on ~1.3% of angles by 1 ulp).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order/size as cv::KeyPoint (28 bytes): pt.x pt.y size angle response octave class_id */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} OrbOracleKeyPoint;

enum { ORB_ORACLE_TIES_CANONICAL = 0, ORB_ORACLE_TIES_NTH_ELEMENT = 1 };
enum { ORB_ORACLE_TRIG_RN = 0, ORB_ORACLE_TRIG_LIBMF = 1 };
enum { ORB_ORACLE_HARRIS_SCORE = 0, ORB_ORACLE_FAST_SCORE = 1 };

#define ORB_ORACLE_MAX_LEVELS 32
#define ORB_ORACLE_EDGE 16 /* EDGE_THRESHOLD, ORBextractor.cc:77 */

/* ctor tables, ORBextractor.cc:457-511 */
typedef struct {
    int nfeatures, nlevels, score_type, fast_th;
    double scale_factor;                         /* double member initialised from a float arg */
    float scale[ORB_ORACLE_MAX_LEVELS];          /* mvScaleFactor */
    float inv_scale[ORB_ORACLE_MAX_LEVELS];      /* mvInvScaleFactor */
    int quota[ORB_ORACLE_MAX_LEVELS];            /* mnFeaturesPerLevel */
    int umax[16];
    int ties_mode, trig_mode;
} OrbOracleParams;

int orb_oracle_params_init(OrbOracleParams *p, int nfeatures, float scale_factor, int nlevels,
                           int score_type, int fast_th);

/* level geometry: ORBextractor.cc:785-786 */
void orb_oracle_level_size(const OrbOracleParams *p, int level, int W, int H, int *w, int *h);

/* per-level cell grid: ORBextractor.cc:527-547.  Returns 0 or a negative error for degenerate grids. */
typedef struct {
    int cols, rows, cell_w, cell_h, n_cells, nf_cell, Wd, Hd;
} OrbOracleCellGrid;
int orb_oracle_cell_grid(const OrbOracleParams *p, int level, int W0, int H0, int w, int h,
                         OrbOracleCellGrid *g);

/* ---- OpenCV primitives (restated; pinned against cv2 by tests/golden) ---- */
/* cv::resize(..., INTER_LINEAR) for CV_8UC1 */
void orb_oracle_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstride,
                                 uint8_t *dst, int dw, int dh, size_t dstride);
/* cv::copyMakeBorder(..., BORDER_REFLECT_101): fills the b-pixel frame around the w x h interior that
 * already sits at buf + b*stride + b */
void orb_oracle_reflect101_frame(uint8_t *buf, int w, int h, size_t stride, int b);
/* FAST-9/16 "m" value of one pixel: max over the 16 contiguous 9-arcs of min(v-ring) / min(ring-v);
 * corner at threshold t <=> m > t; OpenCV score == m-1.  p must have 3 valid pixels all around. */
int orb_oracle_fast_m(const uint8_t *p, size_t stride);
/* cv::FAST(img, kps, th, nonmax=true) on a w x h image: raster-ordered (x,y,score); returns count
 * (stops storing at cap but keeps counting). */
int orb_oracle_fast_detect(const uint8_t *img, int w, int h, size_t stride, int th,
                           int *xs, int *ys, int *scores, int cap);
/* cv::GaussianBlur(roi, roi, Size(7,7), 2, 2, BORDER_REFLECT_101), OpenCV-2.4 integer engine.
 * src is the interior pointer of a buffer that has (at least) a 3-px valid frame around it. */
void orb_oracle_blur7_u8(const uint8_t *src, int w, int h, size_t sstride, uint8_t *dst, size_t dstride);
/* cv::fastAtan2 (degrees, [0,360)) */
float orb_oracle_fast_atan2(float y, float x);
/* ---- reference statics ---- */
/* IC_Angle, ORBextractor.cc:124-151 (center = pixel (x,y) of an image with >=15 valid px around) */
float orb_oracle_ic_angle(const uint8_t *center, size_t stride, const int *umax);
void orb_oracle_ic_moments(const uint8_t *center, size_t stride, const int *umax, int *m01, int *m10);
/* computeOrbDescriptor, ORBextractor.cc:155-194 */
void orb_oracle_brief(const uint8_t *center, size_t stride, float angle_deg, int trig_mode, uint8_t desc[32]);
/* HarrisResponses for one point, ORBextractor.cc:79-120 (blockSize 7) */
float orb_oracle_harris(const uint8_t *img, size_t stride, int x, int y, int block, float k);

/* ---- whole extractor: ORBextractor::operator(), ORBextractor.cc:718-779 ---- */
typedef struct {
    /* optional stage dumps (malloc'd by the oracle, freed by orb_oracle_dump_free); NULL-able request */
    int nlevels;
    int w[ORB_ORACLE_MAX_LEVELS], h[ORB_ORACLE_MAX_LEVELS];
    size_t stride[ORB_ORACLE_MAX_LEVELS];           /* stride of the bordered buffers */
    uint8_t *level[ORB_ORACLE_MAX_LEVELS];          /* bordered, unblurred; interior at +16*stride+16 */
    uint8_t *blurred[ORB_ORACLE_MAX_LEVELS];        /* bordered: interior blurred, frame unblurred */
    int n_level_kp[ORB_ORACLE_MAX_LEVELS];
    long n_ties_at_cut;                             /* how many retainBest cuts fell inside a tie group */
    long n_fallback_cells;                          /* cells that fell back to threshold 7 */
} OrbOracleDump;
void orb_oracle_dump_free(OrbOracleDump *d);

/* returns 0 ok; -1 bad args; -2 degenerate grid; -3 cap too small (n_out still set) */
int orb_oracle_extract(const OrbOracleParams *p, const uint8_t *img, int W, int H, size_t stride,
                       OrbOracleKeyPoint *kps, uint8_t *desc, int cap, int *n_out, OrbOracleDump *dump);

/* retainBest with std::nth_element (orb_oracle_nth.cpp); keeps the first n after introselect */
void orb_oracle_retain_best_nth(OrbOracleKeyPoint *kps, int *n_inout, int n_keep);

/* ---- matcher ---- */
/* ORBmatcher::DescriptorDistance, ORBmatcher.cc:1794-1810 */
int orb_oracle_hamming(const uint8_t *a, const uint8_t *b);

#define ORB_ORACLE_GRID_COLS 64 /* Frame.h:36 */
#define ORB_ORACLE_GRID_ROWS 48 /* Frame.h:35 */

/* the slice of Frame that the matchers read */
typedef struct {
    int n;
    const OrbOracleKeyPoint *keys_un;   /* mvKeysUn */
    const uint8_t *desc;                /* N x 32 */
    float min_x, min_y, max_x, max_y;   /* mnMinX.. (Frame.cc:321-350) */
    float grid_inv_w, grid_inv_h;       /* Frame.cc:77-78 */
    int nlevels;
    const float *scale_factors;         /* Frame::mvScaleFactors, Frame.cc:95-103 */
    /* 64x48 grid, CSR, cell id = ix*48+iy, filled by orb_oracle_frame_grid */
    int cell_start[ORB_ORACLE_GRID_COLS * ORB_ORACLE_GRID_ROWS + 1];
    int *cell_items;                    /* n entries, caller-allocated */
} OrbOracleFrame;

/* Frame.cc:116-123 + PosInGrid :267-277 */
/* cv::Mat `R*x + t` for CV_32F 3x3 / 3x1 (OpenCV's small-matrix gemm path: float accumulation); T = 3x4 row-major [R|t] */
void orb_oracle_cv_Rx_plus_t(const float *T, const float *X, float out[3]);
void orb_oracle_frame_grid(OrbOracleFrame *f);
/* Frame::GetFeaturesInArea, Frame.cc:200-265; returns count */
int orb_oracle_features_in_area(const OrbOracleFrame *f, float x, float y, float r, int min_level,
                                int max_level, int *out, int cap);
/* Frame::mvScaleFactors as Frame.cc:95-103 builds them from GetScaleFactor() */
void orb_oracle_frame_scale_factors(float scale_factor_f32, int nlevels, float *out);

/* ComputeThreeMaxima, ORBmatcher.cc:1748-1789 (histogram given as counts) */
void orb_oracle_three_maxima(const int *counts, int L, int *ind1, int *ind2, int *ind3);

/* SearchByProjection(Frame &Cur, const Frame &Last, float th), ORBmatcher.cc:1507-1620.
 * last_has_mp[i]!=0 <=> LastFrame.mvpMapPoints[i]!=NULL; last_world = 3 floats per Last feature.
 * Tcw = 3x4 row-major float [R|t]. cur_mp_inout[i2] = index of the Last feature whose map point was
 * assigned to Cur feature i2, or -1 (entries >=0 on input are "already occupied"). Returns nmatches. */
int orb_oracle_search_by_projection_ff(const OrbOracleFrame *cur, const OrbOracleFrame *last,
                                       const uint8_t *last_has_mp, const uint8_t *last_outlier,
                                       const float *last_world, const float *Tcw,
                                       float fx, float fy, float cx, float cy, float th,
                                       int check_orientation, int *cur_mp_inout);
/* WindowSearch, ORBmatcher.cc:409-516. f1_has_mp: F1.mvpMapPoints[i]!=NULL && !isBad().
 * out_match21[i2] = i1 or -1 (vnMatches21 / vpMapPointMatches2). */
int orb_oracle_window_search(const OrbOracleFrame *f1, const OrbOracleFrame *f2, const uint8_t *f1_has_mp,
                             int window, int min_level, int max_level, float nnratio,
                             int check_orientation, int *out_match21);
/* SearchForInitialization, ORBmatcher.cc:598-713. prev_matched: 2 floats per F1 feature (in/out).
 * out_match12[i1] = i2 or -1. */
int orb_oracle_search_for_initialization(const OrbOracleFrame *f1, const OrbOracleFrame *f2,
                                         float *prev_matched, int window, float nnratio,
                                         int check_orientation, int *out_match12);
/* SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:49-125 (local-map tracking) */
int orb_oracle_search_local_points(const OrbOracleFrame *f, int npts, const uint8_t *in_view, const float *proj_xy,
                                   const int *level, const float *view_cos, const uint8_t *desc, float th,
                                   float nnratio, int *f_mp);
/* SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), ORBmatcher.cc:1622-1746 (relocalisation) */
int orb_oracle_search_by_projection_kf(const OrbOracleFrame *cur, int npts, const uint8_t *valid, const float *world,
                                       const float *min_dist, const uint8_t *desc, const float *kf_angle,
                                       const float *Tcw, float fx, float fy, float cx, float cy, float th, int orb_dist,
                                       int check_orientation, int *cur_mp);
/* SearchByProjection(Frame &F1, Frame &F2, windowSize, matches2), ORBmatcher.cc:519-594 */
int orb_oracle_search_by_projection_f1f2(const OrbOracleFrame *f1, const OrbOracleFrame *f2, const uint8_t *valid1,
                                         const float *world1, const float *Tc2w, float fx, float fy, float cx, float cy,
                                         int window, float nnratio, int *f2_mp);
/* the guided-search skeleton on explicit queries (rules: see orb_oracle_match.c) */
int orb_oracle_guided_search(const OrbOracleFrame *f, int nq, const float *qu, const float *qv, const float *qr,
                             const int *qlo, const int *qhi, const uint8_t *qdesc, const float *qangle, int rule,
                             float nnratio, int th_dist, int hist_mode, int *slot_owner);
void orb_oracle_guided_best(const OrbOracleFrame *f, int nq, const float *qu, const float *qv, const float *qr,
                            const int *qlo, const int *qhi, const uint8_t *qdesc, int th_dist, int *best_idx);
/* SearchByBoW, ORBmatcher.cc:155-284 (variant 0: KeyFrame vs Frame) and :715-850 (variant 1: KeyFrame vs KeyFrame) */
int orb_oracle_search_by_bow(int variant, int n1, const uint8_t *desc1, const uint8_t *valid1, const float *angle1,
                             int nn1, const int *ids1, const int *ptr1, const int *items1,
                             int n2, const uint8_t *desc2, const uint8_t *valid2, const float *angle2,
                             int nn2, const int *ids2, const int *ptr2, const int *items2,
                             float nnratio, int check_orientation, int *out);
/* SearchForTriangulation + CheckDistEpipolarLine, ORBmatcher.cc:852-1014, :136-153 */
int orb_oracle_search_for_triangulation(int n1, const OrbOracleKeyPoint *keys1, const uint8_t *desc1, const uint8_t *has_mp1,
                                        int nn1, const int *ids1, const int *ptr1, const int *items1,
                                        int n2, const OrbOracleKeyPoint *keys2, const uint8_t *desc2, const uint8_t *has_mp2,
                                        int nn2, const int *ids2, const int *ptr2, const int *items2,
                                        const float *F12, const float *sigma2_kf2, int check_orientation, int *match12);
/* brute-force best/second-best of each query against a database (BASELINE config 5 primitive;
 * same strict-< update rule as every best/second loop in ORBmatcher.cc, e.g. :456-466) */
void orb_oracle_knn2(const uint8_t *q, int nq, const uint8_t *db, long ndb,
                     int *best_dist, int *best_idx, int *second_dist);

/* ---- Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:289-350), SURVEY section 8(f) row N1 ---- */
void orb_oracle_undistort_points(const float *pts, int n, float fx, float fy, float cx, float cy, const float *dist, float *out);
void orb_oracle_undistort_keypoints(const OrbOracleKeyPoint *in, int n, float fx, float fy, float cx, float cy, const float *dist,
                                    OrbOracleKeyPoint *out);
void orb_oracle_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float *dist, float *bounds);

/* ---- SURVEY section 8(f) rows N2 / N4 (orb_oracle_bow.c) ---- */
void orb_oracle_bow_descend(const uint8_t *node_desc, const int32_t *child_ptr, const int32_t *children, int depth_L,
                            const uint8_t *desc, int n, int levelsup, int32_t *leaf_out, int32_t *node_out);
void orb_oracle_bow_transform(const uint8_t *node_desc, const int32_t *child_ptr, const int32_t *children, const int32_t *word_id,
                              const double *weight, int depth_L, int weighting, int norm, const uint8_t *desc, int n,
                              int levelsup, int *nwords_out, int32_t *bow_ids, double *bow_vals, int *nnodes_out,
                              int32_t *fv_ids, int32_t *fv_ptr, int32_t *fv_feats);
void orb_oracle_distinctive(const uint8_t *desc, const int32_t *group_ptr, int ngroups, int32_t *best_out);
int orb_oracle_bow_db_detect(int mode, int nq, const int32_t *q_ids, const double *q_vals, int nkf, const int32_t *kf_ptr,
                             const int32_t *db_ids, const double *db_vals, const uint8_t *connected, const int32_t *covis_ptr,
                             const int32_t *covis, float minScore, int32_t *cand_out, int32_t *common_out, float *score_out);

#ifdef __cplusplus
}
#endif
#endif
