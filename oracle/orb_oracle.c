/*
 * orb_oracle.c -- CPU ORACLE (extractor half).  TEST INFRASTRUCTURE ONLY -- see orb_oracle.h.
 *
 * Restates, stage by stage, reference src/ORBextractor.cc (file:line cited at each function) and the
 * un-vendored OpenCV calls it makes.  Compile with -ffp-contract=off (float ops individually rounded).
 */
#include "orb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const int8_t kPattern[1024] = {
#include "../include/orbfe_brief_pattern.inc"
};

/* cvRound: round-half-to-even on the value widened to double (lrint under the default rounding mode) */
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

/* ------------------------------------------------------------------------------------------------
 * ctor tables -- ORBextractor.cc:457-511
 * ---------------------------------------------------------------------------------------------- */
int orb_oracle_params_init(OrbOracleParams *p, int nfeatures, float scale_factor, int nlevels,
                           int score_type, int fast_th) {
    if (!p || nlevels < 1 || nlevels > ORB_ORACLE_MAX_LEVELS || nfeatures < 0) return -1;
    memset(p, 0, sizeof(*p));
    p->nfeatures = nfeatures;
    p->nlevels = nlevels;
    p->score_type = score_type;
    p->fast_th = fast_th;
    p->scale_factor = (double)scale_factor; /* :459, double member <- float arg */
    const double sf = p->scale_factor;

    p->scale[0] = 1.0f; /* :462-465: float * double -> double -> float */
    for (int i = 1; i < nlevels; i++) p->scale[i] = (float)((double)p->scale[i - 1] * sf);

    const float inv = (float)(1.0f / sf); /* :467 */
    p->inv_scale[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) p->inv_scale[i] = p->inv_scale[i - 1] * inv; /* :470-471 */

    /* :476-487 */
    const float factor = (float)(1.0 / sf);
    float nd = (float)nfeatures * (1.0f - factor) / (1.0f - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        p->quota[l] = cv_round((double)nd);
        sum += p->quota[l];
        nd *= factor;
    }
    p->quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    /* :493-510, HALF_PATCH_SIZE = 15 */
    {
        int umax[17];
        memset(umax, 0, sizeof(umax));
        const int hp = 15;
        int v, v0;
        int vmax = cv_floor((double)((float)hp * sqrtf(2.f) / 2 + 1));
        int vmin = cv_ceil((double)((float)hp * sqrtf(2.f) / 2));
        const double hp2 = (double)(hp * hp);
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round(sqrt(hp2 - (double)(v * v)));
        for (v = hp, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        for (v = 0; v < 16; v++) p->umax[v] = umax[v];
    }
    p->ties_mode = ORB_ORACLE_TIES_CANONICAL;
    p->trig_mode = ORB_ORACLE_TRIG_RN;
    return 0;
}

/* ORBextractor.cc:785-786: Size sz(cvRound((float)image.cols*scale), cvRound((float)image.rows*scale)) */
void orb_oracle_level_size(const OrbOracleParams *p, int level, int W, int H, int *w, int *h) {
    const float s = p->inv_scale[level];
    *w = cv_round((double)((float)W * s));
    *h = cv_round((double)((float)H * s));
}

/* ORBextractor.cc:527-547 */
int orb_oracle_cell_grid(const OrbOracleParams *p, int level, int W0, int H0, int w, int h,
                         OrbOracleCellGrid *g) {
    const float ratio = (float)W0 / (float)H0; /* :527 */
    const int nd = p->quota[level];
    const int cols = (int)sqrtf((float)nd / (5.0f * ratio)); /* :533 */
    const int rows = (int)(ratio * (float)cols);             /* :534 */
    const int Wd = (w - ORB_ORACLE_EDGE) - ORB_ORACLE_EDGE; /* :536-542 */
    const int Hd = (h - ORB_ORACLE_EDGE) - ORB_ORACLE_EDGE;
    if (cols < 1 || rows < 1) {
        /* levelCols == 0 (a quota below 5*ratio) or levelRows == 0 (portrait images): the reference's cell
         * vectors are empty, both loops over rows do nothing and the level yields no keypoints (:549-703);
         * the other levels are unaffected.  An empty grid, not an error. */
        g->cols = 0; g->rows = 0; g->Wd = Wd; g->Hd = Hd;
        g->cell_w = 1; g->cell_h = 1; g->n_cells = 0; g->nf_cell = 0;
        return 0;
    }
    if (Wd < 1 || Hd < 1) return -2;
    g->cols = cols;
    g->rows = rows;
    g->Wd = Wd;
    g->Hd = Hd;
    g->cell_w = (int)ceilf((float)Wd / (float)cols); /* :543-544 */
    g->cell_h = (int)ceilf((float)Hd / (float)rows);
    g->n_cells = rows * cols;
    g->nf_cell = (int)ceilf((float)nd / (float)g->n_cells); /* :546-547 */
    /* Supported domain: every non-last cell must end inside the detectable area [16, dim-16).
     * (Otherwise the reference either throws in cv::Mat::colRange or detects inside the 16-px margin;
     * that needs images only a few dozen pixels wide and is rejected by oracle and product alike.) */
    if ((cols - 1) * g->cell_w > Wd || (rows - 1) * g->cell_h > Hd) return -2;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * cv::resize INTER_LINEAR, CV_8UC1 (OpenCV imgproc resize: 11-bit fixed-point coefficients,
 * HResizeLinear<uchar,int,short,2048> + VResizeLinear<uchar,int,short,FixedPtCast<..,22>> semantics)
 * ---------------------------------------------------------------------------------------------- */
static void resize_axis_tab(int dn, int sn, int *ofs, short *c0, short *c1) {
    const double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; d++) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = cv_floor((double)f);
        f -= (float)s;
        ofs[d] = s;
        float w0 = 1.f - f, w1 = f;
        c0[d] = (short)cv_round((double)(w0 * 2048.f));
        c1[d] = (short)cv_round((double)(w1 * 2048.f));
    }
}

void orb_oracle_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstride,
                                 uint8_t *dst, int dw, int dh, size_t dstride) {
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw);
    short *xa0 = (short *)malloc(sizeof(short) * (size_t)dw), *xa1 = (short *)malloc(sizeof(short) * (size_t)dw);
    int *yofs = (int *)malloc(sizeof(int) * (size_t)dh);
    short *yb0 = (short *)malloc(sizeof(short) * (size_t)dh), *yb1 = (short *)malloc(sizeof(short) * (size_t)dh);
    resize_axis_tab(dw, sw, xofs, xa0, xa1);
    resize_axis_tab(dh, sh, yofs, yb0, yb1);
    /* horizontal clamps: sx<0 -> (0, f=0); sx>=sw-1 -> (sw-1, f=0) */
    for (int d = 0; d < dw; d++) {
        if (xofs[d] < 0) { xofs[d] = 0; xa0[d] = 2048; xa1[d] = 0; }
        if (xofs[d] >= sw - 1) { xofs[d] = sw - 1; xa0[d] = 2048; xa1[d] = 0; }
    }
    int *r0 = (int *)malloc(sizeof(int) * (size_t)dw), *r1 = (int *)malloc(sizeof(int) * (size_t)dw);
    int have0 = -1, have1 = -1;
    for (int dy = 0; dy < dh; dy++) {
        /* vertical: rows sy and sy+1, each clipped to [0, sh-1]; the weights are NOT reset at the clip */
        int sy0 = yofs[dy], sy1 = yofs[dy] + 1;
        if (sy0 < 0) sy0 = 0; if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0; if (sy1 > sh - 1) sy1 = sh - 1;
        if (sy0 == have1) { int *t = r0; r0 = r1; r1 = t; have0 = have1; have1 = -1; }
        if (have0 != sy0) {
            const uint8_t *S = src + (size_t)sy0 * sstride;
            for (int d = 0; d < dw; d++) {
                int s = xofs[d];
                int s1 = s + 1 < sw ? s + 1 : s;
                r0[d] = S[s] * xa0[d] + S[s1] * xa1[d];
            }
            have0 = sy0;
        }
        if (have1 != sy1) {
            if (sy1 == sy0) memcpy(r1, r0, sizeof(int) * (size_t)dw);
            else {
                const uint8_t *S = src + (size_t)sy1 * sstride;
                for (int d = 0; d < dw; d++) {
                    int s = xofs[d];
                    int s1 = s + 1 < sw ? s + 1 : s;
                    r1[d] = S[s] * xa0[d] + S[s1] * xa1[d];
                }
            }
            have1 = sy1;
        }
        const int b0 = yb0[dy], b1 = yb1[dy];
        uint8_t *D = dst + (size_t)dy * dstride;
        for (int d = 0; d < dw; d++) {
            int v = (((b0 * (r0[d] >> 4)) >> 16) + ((b1 * (r1[d] >> 4)) >> 16) + 2) >> 2;
            D[d] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(xa0); free(xa1); free(yofs); free(yb0); free(yb1); free(r0); free(r1);
}

/* cv::copyMakeBorder BORDER_REFLECT_101: index -k -> k, n-1+k -> n-1-k */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

void orb_oracle_reflect101_frame(uint8_t *buf, int w, int h, size_t stride, int b) {
    uint8_t *in = buf + (size_t)b * stride + b;
    for (int y = 0; y < h; y++) {
        uint8_t *row = in + (size_t)y * stride;
        for (int k = 1; k <= b; k++) {
            row[-k] = row[reflect101(-k, w)];
            row[w - 1 + k] = row[reflect101(w - 1 + k, w)];
        }
    }
    for (int k = 1; k <= b; k++) {
        memcpy(in + (ptrdiff_t)(-k) * (ptrdiff_t)stride - b, in + (size_t)reflect101(-k, h) * stride - b, (size_t)(w + 2 * b));
        memcpy(in + (size_t)(h - 1 + k) * stride - b, in + (size_t)reflect101(h - 1 + k, h) * stride - b, (size_t)(w + 2 * b));
    }
}

/* ------------------------------------------------------------------------------------------------
 * cv::FAST 9/16 with non-max suppression (OpenCV features2d fast.cpp semantics)
 * ---------------------------------------------------------------------------------------------- */
static const int kRingX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline int m_from_diffs(const int *d /*16 signed v-ring*/) {
    /* max over the 16 circular 9-arcs of min(d) and of min(-d); clamp at 0.
     * A 9-arc is three consecutive 3-arcs: with the ring unrolled to 24 entries, t[k] = min/max(e[k..k+2]) and
     * arc(k) = min/max(t[k], t[k+3], t[k+6]).  Fixed-size loops without index masking: the compiler vectorises them. */
    int e[24], tmn[22], tmx[22];
    for (int k = 0; k < 16; k++) e[k] = d[k];
    for (int k = 0; k < 8; k++) e[16 + k] = d[k];
    for (int k = 0; k < 22; k++) {
        const int a = e[k], b = e[k + 1], c = e[k + 2];
        const int lo = a < b ? a : b, hi = a > b ? a : b;
        tmn[k] = lo < c ? lo : c;
        tmx[k] = hi > c ? hi : c;
    }
    int best = 0;
    for (int k = 0; k < 16; k++) {
        int mn = tmn[k] < tmn[k + 3] ? tmn[k] : tmn[k + 3];
        mn = mn < tmn[k + 6] ? mn : tmn[k + 6];
        int mx = tmx[k] > tmx[k + 3] ? tmx[k] : tmx[k + 3];
        mx = mx > tmx[k + 6] ? mx : tmx[k + 6];
        if (mn > best) best = mn;   /* all 9 darker than centre by >= mn */
        if (-mx > best) best = -mx; /* all 9 brighter */
    }
    return best;
}

int orb_oracle_fast_m(const uint8_t *p, size_t stride) {
    int d[16];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - (int)p[(ptrdiff_t)kRingY[k] * (ptrdiff_t)stride + kRingX[k]];
    return m_from_diffs(d);
}

/* m for the pixels [j0, j0 + n) of one row, n <= FAST_CHUNK: the same definition as m_from_diffs, laid out as
 * arrays over the columns so that every loop is a plain element-wise loop the compiler vectorises (int16 lanes).
 * Used by orb_oracle_fast_detect; m_from_diffs (behind orb_oracle_fast_m) stays as the per-pixel statement of the same
 * quantity.  Both are pinned against cv2 (tests/test_oracle_golden.py: FAST lists of whole images and ROIs; the m
 * definition against cv2's keypoint responses). */
#define FAST_CHUNK 256
static void fast_m_columns(const uint8_t *row, const ptrdiff_t *ofs, int n, uint8_t *m_out) {
    /* two unsigned passes: dk[k] = max(v - ring_k, 0) ("darker than the centre by"), br[k] = max(ring_k - v, 0).
     * min over an arc of the clamped values = max(min over the arc of the signed values, 0), and m is clamped at 0
     * anyway, so m = max over arcs and polarities of the arc minimum of the clamped values (uint8 lanes). */
    uint8_t dk[16][FAST_CHUNK], br[16][FAST_CHUNK], t0[22][FAST_CHUNK], t1[22][FAST_CHUNK];
    for (int k = 0; k < 16; k++) {
        const uint8_t *r = row + ofs[k];
        for (int x = 0; x < n; x++) {
            const uint8_t v = row[x], q = r[x];
            dk[k][x] = (uint8_t)((v > q ? v : q) - q);
            br[k][x] = (uint8_t)((v > q ? v : q) - v);
        }
    }
    for (int k = 0; k < 22; k++) {
        const uint8_t *a = dk[k & 15], *b = dk[(k + 1) & 15], *c = dk[(k + 2) & 15];
        const uint8_t *A = br[k & 15], *B = br[(k + 1) & 15], *C = br[(k + 2) & 15];
        for (int x = 0; x < n; x++) {
            const uint8_t lo = a[x] < b[x] ? a[x] : b[x], LO = A[x] < B[x] ? A[x] : B[x];
            t0[k][x] = lo < c[x] ? lo : c[x];
            t1[k][x] = LO < C[x] ? LO : C[x];
        }
    }
    for (int x = 0; x < n; x++) m_out[x] = 0;
    for (int k = 0; k < 16; k++) {
        const uint8_t *a = t0[k], *b = t0[k + 3], *c = t0[k + 6], *A = t1[k], *B = t1[k + 3], *C = t1[k + 6];
        for (int x = 0; x < n; x++) {
            uint8_t mn = a[x] < b[x] ? a[x] : b[x];
            mn = mn < c[x] ? mn : c[x];
            uint8_t MN = A[x] < B[x] ? A[x] : B[x];
            MN = MN < C[x] ? MN : C[x];
            uint8_t bb = m_out[x];
            bb = mn > bb ? mn : bb;
            bb = MN > bb ? MN : bb;
            m_out[x] = bb;
        }
    }
}

int orb_oracle_fast_detect(const uint8_t *img, int w, int h, size_t stride, int th,
                           int *xs, int *ys, int *scores, int cap) {
    if (w < 7 || h < 7) return 0;
    ptrdiff_t ofs[16];
    for (int k = 0; k < 16; k++) ofs[k] = (ptrdiff_t)kRingY[k] * (ptrdiff_t)stride + kRingX[k];
    /* three rolling score rows (prev-prev, prev, curr), zero outside [3,w-3) x [3,h-3) */
    uint8_t *buf = (uint8_t *)calloc((size_t)w * 3, 1);
    int n = 0;
    for (int i = 3; i < h - 2; i++) {
        uint8_t *curr = buf + (size_t)((i - 3) % 3) * w;
        memset(curr, 0, (size_t)w);
        if (i < h - 3) {
            const uint8_t *row = img + (size_t)i * stride;
            for (int j0 = 3; j0 < w - 3; j0 += FAST_CHUNK) {
                const int n_ = (w - 3 - j0) < FAST_CHUNK ? (w - 3 - j0) : FAST_CHUNK;
                uint8_t mrow[FAST_CHUNK];
                fast_m_columns(row + j0, ofs, n_, mrow);
                for (int x = 0; x < n_; x++)
                    if (mrow[x] > th) curr[j0 + x] = (uint8_t)(mrow[x] - 1);   /* corner at th <=> m > th; score = m - 1 */
            }
        }
        if (i == 3) continue;
        const uint8_t *prev = buf + (size_t)((i - 4) % 3) * w;
        const uint8_t *pprev = buf + (size_t)((i - 5 + 3) % 3) * w;
        for (int j = 3; j < w - 3; j++) {
            int s = prev[j];
            if (!s) continue;
            if (s > prev[j - 1] && s > prev[j + 1] && s > pprev[j - 1] && s > pprev[j] && s > pprev[j + 1] &&
                s > curr[j - 1] && s > curr[j] && s > curr[j + 1]) {
                if (n < cap) { xs[n] = j; ys[n] = i - 1; scores[n] = s; }
                n++;
            }
        }
    }
    free(buf);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * cv::GaussianBlur 7x7 sigma 2, u8, BORDER_REFLECT_101 -- OpenCV-2.4 integer engine:
 * getGaussianKernel(7,2) -> *256 -> [18,34,49,55,49,34,18]; int32 row pass, int32 column pass,
 * result = round-half-even(sum / 65536), saturated.
 * ---------------------------------------------------------------------------------------------- */
void orb_oracle_blur7_u8(const uint8_t *src, int w, int h, size_t sstride, uint8_t *dst, size_t dstride) {
    static const int K[4] = {55, 49, 34, 18}; /* centre, +-1, +-2, +-3 */
    int *rows = (int *)malloc(sizeof(int) * (size_t)w * 7);
    /* ring of 7 row-filtered lines for source rows y-3..y+3 (taken from the real frame, which holds
     * the reflect-101 pixels of the unblurred image) */
    for (int sy = -3; sy < h + 3; sy++) {
        int *R = rows + (size_t)((sy + 3) % 7) * w;
        const uint8_t *S = src + (ptrdiff_t)sy * (ptrdiff_t)sstride;
        for (int x = 0; x < w; x++) {
            R[x] = K[0] * S[x] + K[1] * (S[x - 1] + S[x + 1]) + K[2] * (S[x - 2] + S[x + 2]) +
                   K[3] * (S[x - 3] + S[x + 3]);
        }
        int y = sy - 3;
        if (y < 0) continue;
        const int *r0 = rows + (size_t)((y + 0) % 7) * w; /* source row y-3 */
        const int *r1 = rows + (size_t)((y + 1) % 7) * w;
        const int *r2 = rows + (size_t)((y + 2) % 7) * w;
        const int *r3 = rows + (size_t)((y + 3) % 7) * w; /* centre */
        const int *r4 = rows + (size_t)((y + 4) % 7) * w;
        const int *r5 = rows + (size_t)((y + 5) % 7) * w;
        const int *r6 = rows + (size_t)((y + 6) % 7) * w;
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            int s = K[0] * r3[x] + K[1] * (r2[x] + r4[x]) + K[2] * (r1[x] + r5[x]) + K[3] * (r0[x] + r6[x]);
            int q = s >> 16, r = s & 0xFFFF;
            q += (r > 0x8000) | ((r == 0x8000) & (q & 1));
            D[x] = (uint8_t)(q > 255 ? 255 : q);
        }
    }
    free(rows);
}

/* ------------------------------------------------------------------------------------------------
 * cv::fastAtan2 -- 7th-order odd polynomial, all binary32, no FMA
 * ---------------------------------------------------------------------------------------------- */
float orb_oracle_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* IC_Angle -- ORBextractor.cc:124-151 */
void orb_oracle_ic_moments(const uint8_t *center, size_t stride, const int *umax, int *m01o, int *m10o) {
    int m_01 = 0, m_10 = 0;
    for (int u = -15; u <= 15; ++u) m_10 += u * center[u]; /* :131-132 */
    const ptrdiff_t step = (ptrdiff_t)stride;
    for (int v = 1; v <= 15; ++v) { /* :136-148 */
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    *m01o = m_01;
    *m10o = m_10;
}

float orb_oracle_ic_angle(const uint8_t *center, size_t stride, const int *umax) {
    int m01, m10;
    orb_oracle_ic_moments(center, stride, umax, &m01, &m10);
    return orb_oracle_fast_atan2((float)m01, (float)m10); /* :150 */
}

/* computeOrbDescriptor -- ORBextractor.cc:155-194 */
void orb_oracle_brief(const uint8_t *center, size_t stride, float angle_deg, int trig_mode, uint8_t desc[32]) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f); /* :154 */
    const float angle = angle_deg * factorPI;                       /* :159 */
    float a, b;
    if (trig_mode == ORB_ORACLE_TRIG_LIBMF) { a = cosf(angle); b = sinf(angle); }
    else { a = (float)cos((double)angle); b = (float)sin((double)angle); } /* :160 */
    const ptrdiff_t step = (ptrdiff_t)stride;
    const int8_t *pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) { /* 16 points = 8 pairs per byte */
        int val = 0;
        for (int k = 0; k < 8; k++) {
            const float x0 = (float)pat[4 * k + 0], y0 = (float)pat[4 * k + 1];
            const float x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
            /* :165-167: center[cvRound(x*b + y*a)*step + cvRound(x*a - y*b)] */
            int t0 = center[cv_round((double)(x0 * b + y0 * a)) * step + cv_round((double)(x0 * a - y0 * b))];
            int t1 = center[cv_round((double)(x1 * b + y1 * a)) * step + cv_round((double)(x1 * a - y1 * b))];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* HarrisResponses -- ORBextractor.cc:79-120 for one point (pt already in img coordinates) */
float orb_oracle_harris(const uint8_t *img, size_t stride, int x, int y, int block, float harris_k) {
    const ptrdiff_t step = (ptrdiff_t)stride;
    const int r = block / 2;
    float scale = (float)(1 << 2) * (float)block * 255.0f;
    scale = 1.0f / scale;
    const float scale_sq_sq = scale * scale * scale * scale;
    const uint8_t *ptr0 = img + (ptrdiff_t)(y - r) * step + (x - r);
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < block; i++)
        for (int j = 0; j < block; j++) {
            const uint8_t *ptr = ptr0 + i * step + j;
            int Ix = (ptr[1] - ptr[-1]) * 2 + (ptr[-step + 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[step - 1]);
            int Iy = (ptr[step] - ptr[-step]) * 2 + (ptr[step - 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[-step + 1]);
            a += Ix * Ix;
            b += Iy * Iy;
            c += Ix * Iy;
        }
    return ((float)a * (float)b - (float)c * (float)c - harris_k * ((float)a + (float)b) * ((float)a + (float)b)) * scale_sq_sq;
}

/* ------------------------------------------------------------------------------------------------
 * KeyPointsFilter::retainBest + the reference's resize(n)  (ORBextractor.cc:683-685, :697-701)
 * canonical rule: keep the n largest responses; ties at the cut -> earlier position in the vector wins;
 * survivors keep their original relative order.
 * ---------------------------------------------------------------------------------------------- */
static int cmp_desc_f32(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x < y) - (x > y);
}
static long g_ties_at_cut; /* diagnostic only (single-threaded use in tests) */

static void retain_best_canonical(OrbOracleKeyPoint *k, int *n_io, int keep) {
    int n = *n_io;
    if (keep < 0) keep = 0;
    if (n <= keep) return;
    if (keep == 0) { *n_io = 0; return; }
    /* responses are small non-negative integers for FAST_SCORE; generic path sorts a copy */
    float *r = (float *)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; i++) r[i] = k[i].response;
    /* find the keep-th largest value */
    float *s = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(s, r, sizeof(float) * (size_t)n);
    qsort(s, (size_t)n, sizeof(float), cmp_desc_f32);
    const float cut = s[keep - 1];
    int n_gt = 0, n_eq = 0;
    for (int i = 0; i < n; i++) { n_gt += r[i] > cut; n_eq += r[i] == cut; }
    int take_eq = keep - n_gt;
    if (take_eq < n_eq) g_ties_at_cut++;
    int o = 0;
    for (int i = 0; i < n; i++) {
        if (r[i] > cut) k[o++] = k[i];
        else if (r[i] == cut && take_eq > 0) { k[o++] = k[i]; take_eq--; }
    }
    *n_io = o;
    free(r);
    free(s);
}

static void retain_best(const OrbOracleParams *p, OrbOracleKeyPoint *k, int *n_io, int keep) {
    if (p->ties_mode == ORB_ORACLE_TIES_NTH_ELEMENT) orb_oracle_retain_best_nth(k, n_io, keep);
    else retain_best_canonical(k, n_io, keep);
}

void orb_oracle_dump_free(OrbOracleDump *d) {
    if (!d) return;
    for (int l = 0; l < ORB_ORACLE_MAX_LEVELS; l++) {
        free(d->level[l]); d->level[l] = NULL;
        free(d->blurred[l]); d->blurred[l] = NULL;
    }
}

/* ------------------------------------------------------------------------------------------------
 * ORBextractor::operator() -- ORBextractor.cc:718-779
 * ---------------------------------------------------------------------------------------------- */
int orb_oracle_extract(const OrbOracleParams *p, const uint8_t *img, int W, int H, size_t stride,
                       OrbOracleKeyPoint *kps_out, uint8_t *desc_out, int cap, int *n_out, OrbOracleDump *dump) {
    if (!p || !n_out) return -1;
    *n_out = 0;
    if (!img || W <= 0 || H <= 0) return 0; /* :721-722 empty image -> silent return */
    const int L = p->nlevels;
    const int E = ORB_ORACLE_EDGE;
    int lw[ORB_ORACLE_MAX_LEVELS], lh[ORB_ORACLE_MAX_LEVELS];
    size_t ls[ORB_ORACLE_MAX_LEVELS];
    uint8_t *lev[ORB_ORACLE_MAX_LEVELS];
    OrbOracleCellGrid grid[ORB_ORACLE_MAX_LEVELS];
    memset(lev, 0, sizeof(lev));
    int rc = 0;
    long n_fallback = 0;
    g_ties_at_cut = 0;

    /* ---- ComputePyramid, :781-822 ---- */
    for (int l = 0; l < L; l++) {
        orb_oracle_level_size(p, l, W, H, &lw[l], &lh[l]);
        if (lw[l] < 1 || lh[l] < 1) { rc = -2; goto done; }
        ls[l] = (size_t)(lw[l] + 2 * E);
        lev[l] = (uint8_t *)malloc(ls[l] * (size_t)(lh[l] + 2 * E));
        uint8_t *in = lev[l] + (size_t)E * ls[l] + E;
        if (l == 0) {
            for (int y = 0; y < H; y++) memcpy(in + (size_t)y * ls[0], img + (size_t)y * stride, (size_t)W);
        } else {
            const uint8_t *pin = lev[l - 1] + (size_t)E * ls[l - 1] + E;
            orb_oracle_resize_linear_u8(pin, lw[l - 1], lh[l - 1], ls[l - 1], in, lw[l], lh[l], ls[l]); /* :800 */
        }
        orb_oracle_reflect101_frame(lev[l], lw[l], lh[l], ls[l], E); /* :806 / :814 */
    }
    for (int l = 0; l < L; l++) {
        int g = orb_oracle_cell_grid(p, l, lw[0], lh[0], lw[l], lh[l], &grid[l]);
        if (g) { rc = -2; goto done; }
    }

    /* ---- ComputeKeyPoints, :522-707 ---- */
    OrbOracleKeyPoint *levkp[ORB_ORACLE_MAX_LEVELS];
    int levn[ORB_ORACLE_MAX_LEVELS];
    memset(levkp, 0, sizeof(levkp));
    for (int l = 0; l < L; l++) {
        const OrbOracleCellGrid *g = &grid[l];
        const int nd = p->quota[l];
        const uint8_t *in = lev[l] + (size_t)E * ls[l] + E;
        const int maxBX = lw[l] - E, maxBY = lh[l] - E;
        const int nCells = g->n_cells;
        /* per-cell candidate storage */
        const int ccap = ((g->cell_w + 1) / 2 + 1) * ((g->cell_h + 1) / 2 + 1) + 8;
        int *cx = (int *)malloc(sizeof(int) * (size_t)ccap), *cy = (int *)malloc(sizeof(int) * (size_t)ccap),
            *cs = (int *)malloc(sizeof(int) * (size_t)ccap);
        OrbOracleKeyPoint **cell = (OrbOracleKeyPoint **)calloc((size_t)nCells, sizeof(*cell));
        int *nTotal = (int *)calloc((size_t)nCells, sizeof(int));
        int *nRetain = (int *)calloc((size_t)nCells, sizeof(int));
        uint8_t *noMore = (uint8_t *)calloc((size_t)nCells, 1);
        int *iniXc = (int *)calloc((size_t)g->cols, sizeof(int)), *iniYr = (int *)calloc((size_t)g->rows, sizeof(int));
        int nNoMore = 0, nToDist = 0;

        float hY = (float)(g->cell_h + 6); /* :557 */
        for (int i = 0; i < g->rows; i++) {
            const float iniY = (float)(E + i * g->cell_h - 3); /* :561 */
            iniYr[i] = (int)iniY;
            if (i == g->rows - 1) {
                hY = (float)(maxBY + 3) - iniY; /* :566 */
                if (hY <= 0) continue;
            }
            float hX = (float)(g->cell_w + 6); /* :571 */
            for (int j = 0; j < g->cols; j++) {
                float iniX = (float)(E + j * g->cell_w - 3); /* :579-586 (same value for every i) */
                iniXc[j] = (int)iniX;
                if (j == g->cols - 1) {
                    hX = (float)(maxBX + 3) - iniX; /* :592 */
                    if (hX <= 0) continue;
                }
                const int cwid = (int)hX, chei = (int)hY;
                const uint8_t *cimg = in + (ptrdiff_t)iniYr[i] * (ptrdiff_t)ls[l] + iniXc[j];
                int n = orb_oracle_fast_detect(cimg, cwid, chei, ls[l], p->fast_th, cx, cy, cs, ccap); /* :607 */
                if (n <= 3) { /* :609-614 */
                    n = orb_oracle_fast_detect(cimg, cwid, chei, ls[l], 7, cx, cy, cs, ccap);
                    n_fallback++;
                }
                if (n > ccap) { rc = -1; n = ccap; }
                const int ci = i * g->cols + j;
                cell[ci] = (OrbOracleKeyPoint *)malloc(sizeof(OrbOracleKeyPoint) * (size_t)(n > 0 ? n : 1));
                for (int k = 0; k < n; k++) {
                    OrbOracleKeyPoint kp = {(float)cx[k], (float)cy[k], 7.f, -1.f, (float)cs[k], 0, -1};
                    if (p->score_type == ORB_ORACLE_HARRIS_SCORE) /* :616-620 */
                        kp.response = orb_oracle_harris(cimg, ls[l], cx[k], cy[k], 7, 0.04f);
                    cell[ci][k] = kp;
                }
                nTotal[ci] = n; /* :622-636 */
                if (n > g->nf_cell) { nRetain[ci] = g->nf_cell; noMore[ci] = 0; }
                else { nRetain[ci] = n; nToDist += g->nf_cell - n; noMore[ci] = 1; nNoMore++; }
            }
        }
        /* :644-670 */
        while (nToDist > 0 && nNoMore < nCells) {
            int nNew = g->nf_cell + (int)ceilf((float)nToDist / (float)(nCells - nNoMore));
            nToDist = 0;
            for (int ci = 0; ci < nCells; ci++) {
                if (noMore[ci]) continue;
                if (nTotal[ci] > nNew) { nRetain[ci] = nNew; noMore[ci] = 0; }
                else { nRetain[ci] = nTotal[ci]; nToDist += nNew - nTotal[ci]; noMore[ci] = 1; nNoMore++; }
            }
        }
        /* :672-695 */
        levkp[l] = (OrbOracleKeyPoint *)malloc(sizeof(OrbOracleKeyPoint) * (size_t)(nd * 2 + nCells * 4 + 64));
        int ln = 0, lcap = nd * 2 + nCells * 4 + 64;
        const float scaledPatch = (float)(int)(31.0f * p->scale[l]); /* :675 */
        for (int i = 0; i < g->rows; i++)
            for (int j = 0; j < g->cols; j++) {
                const int ci = i * g->cols + j;
                int n = nTotal[ci];
                if (!cell[ci]) continue;
                retain_best(p, cell[ci], &n, nRetain[ci]); /* :683 */
                if (n > nRetain[ci]) n = nRetain[ci];      /* :684-685 */
                if (ln + n > lcap) {
                    lcap = (ln + n) * 2;
                    levkp[l] = (OrbOracleKeyPoint *)realloc(levkp[l], sizeof(OrbOracleKeyPoint) * (size_t)lcap);
                }
                for (int k = 0; k < n; k++) { /* :687-694 */
                    OrbOracleKeyPoint kp = cell[ci][k];
                    kp.x += (float)iniXc[j];
                    kp.y += (float)iniYr[i];
                    kp.octave = l;
                    kp.size = scaledPatch;
                    levkp[l][ln++] = kp;
                }
            }
        if (ln > nd) { /* :697-701 */
            retain_best(p, levkp[l], &ln, nd);
            if (ln > nd) ln = nd;
        }
        levn[l] = ln;
        for (int ci = 0; ci < nCells; ci++) free(cell[ci]);
        free(cell); free(nTotal); free(nRetain); free(noMore); free(iniXc); free(iniYr);
        free(cx); free(cy); free(cs);
    }
    /* :705-706 orientations on the unblurred pyramid */
    for (int l = 0; l < L; l++) {
        const uint8_t *in = lev[l] + (size_t)E * ls[l] + E;
        for (int k = 0; k < levn[l]; k++) {
            OrbOracleKeyPoint *kp = &levkp[l][k];
            kp->angle = orb_oracle_ic_angle(in + (size_t)cv_round((double)kp->y) * ls[l] + cv_round((double)kp->x), ls[l], p->umax);
        }
    }

    /* ---- descriptors, :749-778 ---- */
    int total = 0;
    for (int l = 0; l < L; l++) total += levn[l];
    *n_out = total;
    if (dump) {
        memset(dump, 0, sizeof(*dump));
        dump->nlevels = L;
        dump->n_fallback_cells = n_fallback;
    }
    int off = 0;
    for (int l = 0; l < L; l++) {
        uint8_t *bl = NULL;
        if (levn[l] > 0 || dump) {
            bl = (uint8_t *)malloc(ls[l] * (size_t)(lh[l] + 2 * E));
            memcpy(bl, lev[l], ls[l] * (size_t)(lh[l] + 2 * E)); /* frame stays unblurred */
            orb_oracle_blur7_u8(lev[l] + (size_t)E * ls[l] + E, lw[l], lh[l], ls[l], bl + (size_t)E * ls[l] + E, ls[l]); /* :760 */
        }
        for (int k = 0; k < levn[l]; k++) {
            if (off + k >= cap) { rc = rc ? rc : -3; break; }
            OrbOracleKeyPoint kp = levkp[l][k];
            const uint8_t *c = bl + (size_t)E * ls[l] + E + (size_t)cv_round((double)kp.y) * ls[l] + cv_round((double)kp.x);
            if (desc_out) orb_oracle_brief(c, ls[l], kp.angle, p->trig_mode, desc_out + (size_t)(off + k) * 32); /* :764 */
            if (l != 0) { /* :768-775 */
                const float s = p->scale[l];
                kp.x *= s;
                kp.y *= s;
            }
            if (kps_out) kps_out[off + k] = kp;
        }
        off += levn[l];
        if (dump) {
            dump->w[l] = lw[l]; dump->h[l] = lh[l]; dump->stride[l] = ls[l];
            dump->level[l] = lev[l]; lev[l] = NULL;
            dump->blurred[l] = bl; bl = NULL;
            dump->n_level_kp[l] = levn[l];
        }
        free(bl);
    }
    if (dump) dump->n_ties_at_cut = g_ties_at_cut;
    for (int l = 0; l < L; l++) free(levkp[l]);
done:
    for (int l = 0; l < L; l++) free(lev[l]);
    return rc;
}
