// e2e_driver.cpp -- bench harness (NOT part of the product library): the end-to-end stream pipeline of bench.py in C++.
//
// What a C++ host (the reference's Tracking thread, src/Tracking.cc:215-236 + :565) does per frame -- construct a Frame
// from a host image (ORBextractor::operator(), Frame.cc:60), then ORBmatcher::SearchByProjection(CurrentFrame, LastFrame)
// -- is done here for a stream of frames through the PUBLIC C-ABI of liborbfe.so only, on plain std::threads:
//   * `nex` extractor handles alternate batches (orbfe_extract_batch: pinned host frames in, host keypoints / descriptors
//     out), so the upload of batch t+1 overlaps the kernels of batch t;
//   * `nmatch` matcher handles alternate batches: they build the frame views (the slice of Frame the matcher reads) and the synthetic
//     map points (every keypoint back-projected at a fixed depth, same float32 operations as bench.py's backproject())
//     and orbfe_search_by_projection_frames on HOST views (its H2D / D2H inside).
// bench.py used to run this pipeline on Python threads; the GIL hand-offs between seven threads cost more than the
// library calls' own host time, which says nothing about the library.  No CUDA call is made here.
//
// Build: g++ -O2 -shared -fPIC -std=c++17 -I include tools/e2e_driver.cpp -o orb_slam_b200/libe2e_driver.so
//        (orb_slam_b200/build.py does it; liborbfe.so is resolved at load time through the handle bench.py passes in).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "orbfe.h"
#include "orbfe_match.h"

namespace {

struct Api {  // resolved from liborbfe.so by the caller (dlsym through ctypes): the driver has no link-time dependency
    int (*extractor_create)(int, float, int, int, int, int, OrbfeExtractor **);
    int (*extractor_destroy)(OrbfeExtractor *);
    int (*extract_batch)(OrbfeExtractor *, const uint8_t *, int, int, size_t, size_t, int, OrbfeKeyPoint *, uint8_t *, int, int *);
    int (*extractor_last_launches)(const OrbfeExtractor *);
    int (*matcher_create)(int, OrbfeMatcher **);
    int (*matcher_destroy)(OrbfeMatcher *);
    int (*matcher_counters)(const OrbfeMatcher *, unsigned long long *, unsigned long long *, unsigned long long *);
    int (*sbp_frames)(OrbfeMatcher *, int, const OrbfeFrameView *, const OrbfeFrameView *, const uint8_t *const *, const uint8_t *const *,
                      const float *const *, const float *const *, float, float, float, float, float, int, int *const *, int *);
    void (*frame_scale_factors)(float, int, float *);
    const char *(*last_error)(void);
};

struct Tail {  // private copy of a batch's last frame: the Last frame of the next batch's first pair
    std::vector<OrbfeKeyPoint> kps;
    std::vector<uint8_t> desc;
    std::vector<float> world;
    int n = 0;
};

struct Driver {
    Api api;
    int W, H, nfeat, nlevels, B, NB, nex, nmatch, nbuf, device;
    float scale, fx, fy, cx, cy, depth, th;
    const uint8_t *frames;     // pinned host, NB * B frames
    const float *Tcws;         // [NB * B][12]
    std::vector<OrbfeKeyPoint *> kps;   // nbuf pinned output sets
    std::vector<uint8_t *> desc;
    std::vector<int *> cnt;
    std::vector<OrbfeExtractor *> ex;
    std::vector<OrbfeMatcher *> mt;
    std::vector<float> sf;
    std::vector<uint8_t> ones, zeros;
    long long step = 0;        // batches processed so far (continues across runs)
    // results of the last matched batch (for the oracle cross-check)
    long long last_st = -1;
    std::vector<int> last_mp;  // first 4 pairs
    std::string err;
};

struct Run {
    Driver *D;
    long long s0;
    int nsub;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> extracted, tail_ready, matched;   // per batch of this run
    std::vector<Tail> tails;                           // one per batch of the run (sized before the threads start: never reallocated)
    std::atomic<long long> kp{0}, nm{0}, launches{0};
    std::atomic<int> failed{0};
    double t_extract = 0, t_views = 0, t_match = 0;    // summed host seconds inside the calls (under mu)

    void set(std::vector<int> &v, int i) {
        { std::lock_guard<std::mutex> g(mu); v[i] = 1; }
        cv.notify_all();
    }
    void wait(std::vector<int> &v, int i) {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return v[i] != 0 || failed.load() != 0; });
    }
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void extract_thread(Run *R, int id) {
    Driver *D = R->D;
    const size_t fbytes = (size_t)D->W * D->H;
    for (int i = id; i < R->nsub && !R->failed.load(); i += D->nex) {
        const long long st = R->s0 + i;
        if (i >= D->nbuf) R->wait(R->matched, i - D->nbuf);   // output set st % nbuf was last used by batch st - nbuf
        if (R->failed.load()) break;
        const int b = (int)(st % D->nbuf);
        const double t0 = now();
        const int rc = D->api.extract_batch(D->ex[id], D->frames + (size_t)(st % D->NB) * D->B * fbytes, D->W, D->H, D->W, fbytes, D->B,
                                            D->kps[b], D->desc[b], D->nfeat, D->cnt[b]);
        const double dt = now() - t0;
        if (rc != 0) {
            std::lock_guard<std::mutex> g(R->mu);
            D->err = std::string("orbfe_extract_batch: ") + D->api.last_error();
            R->failed = 1;
            R->cv.notify_all();
            break;
        }
        long long k = 0;
        for (int f = 0; f < D->B; f++) k += D->cnt[b][f];
        R->kp += k;
        R->launches += D->api.extractor_last_launches(D->ex[id]);
        { std::lock_guard<std::mutex> g(R->mu); R->t_extract += dt; }
        R->set(R->extracted, i);
    }
}

void match_thread(Run *R, int id) {
    Driver *D = R->D;
    const int B = D->B, cap = D->nfeat;
    std::vector<OrbfeFrameView> cur(B), last(B);
    std::vector<float> world((size_t)B * cap * 3);
    std::vector<const uint8_t *> has(B, D->ones.data()), outl(B, D->zeros.data());
    std::vector<const float *> wptr(B), tptr(B);
    std::vector<int> mp((size_t)B * cap), nmv(B);
    std::vector<int *> mptr(B);
    for (int i = id; i < R->nsub && !R->failed.load(); i += D->nmatch) {
        const long long st = R->s0 + i;
        R->wait(R->extracted, i);
        if (R->failed.load()) break;
        const int b = (int)(st % D->nbuf);
        const double t0 = now();
        const OrbfeKeyPoint *K = D->kps[b];
        const uint8_t *Dd = D->desc[b];
        const int *cn = D->cnt[b];
        for (int f = 0; f < B; f++) {
            OrbfeFrameView &v = cur[f];
            v.n = cn[f];
            v.keys_un = K + (size_t)f * cap;
            v.desc = Dd + (size_t)f * cap * 32;
            v.min_x = 0.f; v.min_y = 0.f; v.max_x = (float)D->W; v.max_y = (float)D->H;      // Frame.cc:342-348 (no distortion)
            v.grid_inv_w = 64.f / (float)D->W;                                                // Frame.cc:77-78
            v.grid_inv_h = 48.f / (float)D->H;
            v.nlevels = D->nlevels;
            v.scale_factors = D->sf.data();
            float *w = &world[(size_t)f * cap * 3];
            const OrbfeKeyPoint *kf = v.keys_un;
            for (int k = 0; k < v.n; k++) {   // every operation rounded to binary32, as numpy does in bench.py's backproject()
                const float a = (kf[k].x - D->cx) / D->fx, c = (kf[k].y - D->cy) / D->fy;
                w[3 * k] = a * D->depth;
                w[3 * k + 1] = c * D->depth;
                w[3 * k + 2] = D->depth;
            }
        }
        // this batch's own tail first: batch st + 1 only needs that, not our matches
        Tail &T = R->tails[i];
        T.n = cn[B - 1];
        T.kps.assign(K + (size_t)(B - 1) * cap, K + (size_t)(B - 1) * cap + T.n);
        T.desc.assign(Dd + (size_t)(B - 1) * cap * 32, Dd + (size_t)(B - 1) * cap * 32 + (size_t)T.n * 32);
        T.world.assign(&world[(size_t)(B - 1) * cap * 3], &world[(size_t)(B - 1) * cap * 3] + (size_t)T.n * 3);
        R->set(R->tail_ready, i);
        const Tail *P = &T;     // the first batch of a run has no predecessor: its own last frame stands in (a scene cut)
        if (i > 0) {
            R->wait(R->tail_ready, i - 1);
            if (R->failed.load()) break;
            P = &R->tails[i - 1];
        }
        for (int f = 0; f < B; f++) {
            if (f == 0) {
                last[0] = cur[B - 1];
                last[0].n = P->n;
                last[0].keys_un = P->kps.data();
                last[0].desc = P->desc.data();
                wptr[0] = P->world.data();
            } else {
                last[f] = cur[f - 1];
                wptr[f] = &world[(size_t)(f - 1) * cap * 3];
            }
            tptr[f] = D->Tcws + ((size_t)(st % D->NB) * B + f) * 12;
            mptr[f] = &mp[(size_t)f * cap];
        }
        std::fill(mp.begin(), mp.end(), -1);
        const double t1 = now();
        const int rc = D->api.sbp_frames(D->mt[id], B, cur.data(), last.data(), has.data(), outl.data(), wptr.data(), tptr.data(), D->fx, D->fy,
                                         D->cx, D->cy, D->th, 1, mptr.data(), nmv.data());
        const double t2 = now();
        if (rc != 0) {
            std::lock_guard<std::mutex> g(R->mu);
            D->err = std::string("orbfe_search_by_projection_frames: ") + D->api.last_error();
            R->failed = 1;
            R->cv.notify_all();
            break;
        }
        long long s = 0;
        for (int f = 0; f < B; f++) s += nmv[f];
        R->nm += s;
        {
            std::lock_guard<std::mutex> g(R->mu);
            R->t_views += t1 - t0;
            R->t_match += t2 - t1;
            if (st > D->last_st) {
                D->last_st = st;
                D->last_mp.assign(mp.begin(), mp.begin() + (size_t)4 * cap);
            }
        }
        R->set(R->matched, i);
    }
}

}  // namespace

extern "C" {

struct E2eConfig {
    int W, H, nfeat, nlevels, fast_th, B, NB, nex, nmatch, nbuf, device;
    float scale, fx, fy, cx, cy, depth, th;
};

// fns: the ten liborbfe.so entry points in the order of struct Api; out_* : nbuf pinned host buffers each
void *e2e_create(const E2eConfig *c, void *const *fns, const uint8_t *frames, const float *Tcws, void *const *out_kps, void *const *out_desc,
                 void *const *out_cnt) {
    Driver *D = new Driver();
    std::memcpy(&D->api, fns, sizeof(Api));
    D->W = c->W; D->H = c->H; D->nfeat = c->nfeat; D->nlevels = c->nlevels; D->B = c->B; D->NB = c->NB;
    D->nex = c->nex; D->nmatch = c->nmatch; D->nbuf = c->nbuf; D->device = c->device;
    D->scale = c->scale; D->fx = c->fx; D->fy = c->fy; D->cx = c->cx; D->cy = c->cy; D->depth = c->depth; D->th = c->th;
    D->frames = frames;
    D->Tcws = Tcws;
    for (int i = 0; i < c->nbuf; i++) {
        D->kps.push_back((OrbfeKeyPoint *)out_kps[i]);
        D->desc.push_back((uint8_t *)out_desc[i]);
        D->cnt.push_back((int *)out_cnt[i]);
    }
    D->sf.resize(c->nlevels);
    D->api.frame_scale_factors(c->scale, c->nlevels, D->sf.data());
    D->ones.assign(c->nfeat, 1);
    D->zeros.assign(c->nfeat, 0);
    for (int i = 0; i < c->nex; i++) {
        OrbfeExtractor *x = nullptr;
        if (D->api.extractor_create(c->nfeat, c->scale, c->nlevels, 1 /* FAST_SCORE */, c->fast_th, c->device, &x) != 0) { delete D; return nullptr; }
        D->ex.push_back(x);
    }
    for (int i = 0; i < c->nmatch; i++) {
        OrbfeMatcher *m = nullptr;
        if (D->api.matcher_create(c->device, &m) != 0) { delete D; return nullptr; }
        D->mt.push_back(m);
    }
    return D;
}

void e2e_destroy(void *h) {
    Driver *D = (Driver *)h;
    if (!D) return;
    for (auto *x : D->ex) D->api.extractor_destroy(x);
    for (auto *m : D->mt) D->api.matcher_destroy(m);
    delete D;
}

// Processes `nsub` batches (continuing the stream where the previous run stopped).  out[0..6] = keypoints, matches, kernel launches of the
// extractors, matcher H2D bytes, matcher D2H bytes, matcher launches (deltas of this run), batch index (st % NB) of the last matched batch;
// host_s[0..2] = summed host seconds inside orbfe_extract_batch, the view/map-point glue, orbfe_search_by_projection_frames.
int e2e_run(void *h, int nsub, long long *out, double *host_s) {
    Driver *D = (Driver *)h;
    Run R;
    R.D = D;
    R.s0 = D->step;
    R.nsub = nsub;
    R.extracted.assign(nsub, 0);
    R.tail_ready.assign(nsub, 0);
    R.matched.assign(nsub, 0);
    R.tails.resize(nsub);
    unsigned long long c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0}, t[3];
    for (auto *m : D->mt) { D->api.matcher_counters(m, &t[0], &t[1], &t[2]); for (int k = 0; k < 3; k++) c0[k] += t[k]; }
    std::vector<std::thread> th;
    for (int i = 0; i < D->nex; i++) th.emplace_back(extract_thread, &R, i);
    for (int i = 0; i < D->nmatch; i++) th.emplace_back(match_thread, &R, i);
    for (auto &x : th) x.join();
    for (auto *m : D->mt) { D->api.matcher_counters(m, &t[0], &t[1], &t[2]); for (int k = 0; k < 3; k++) c1[k] += t[k]; }
    D->step += nsub;
    out[0] = R.kp; out[1] = R.nm; out[2] = R.launches;
    out[3] = (long long)(c1[0] - c0[0]); out[4] = (long long)(c1[1] - c0[1]); out[5] = (long long)(c1[2] - c0[2]);
    out[6] = D->last_st;
    host_s[0] = R.t_extract; host_s[1] = R.t_views; host_s[2] = R.t_match;
    return R.failed.load() ? -1 : 0;
}

// match vectors of the first four pairs of the last matched batch (4 x nfeat ints)
int e2e_last_matches(void *h, int *mp_out) {
    Driver *D = (Driver *)h;
    if (D->last_mp.empty()) return -1;
    std::memcpy(mp_out, D->last_mp.data(), D->last_mp.size() * sizeof(int));
    return 0;
}

const char *e2e_error(void *h) { return ((Driver *)h)->err.c_str(); }

}  // extern "C"
