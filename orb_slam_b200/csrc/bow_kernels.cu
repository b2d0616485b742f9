// bow_kernels.cu -- the two Hamming-primitive rows of SURVEY.md section 8(f):
//   N2  DBoW2 vocabulary-tree descent (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1216-1260, FORB.cpp:79-99)
//   N4  MapPoint::ComputeDistinctiveDescriptors, batched (reference src/MapPoint.cc:185-250)
// plus the device-side half of the C-ABI in include/orbfe_bow.h.  The std::map builds of transform() live in
// host/bow_host.cpp.  Integer work only (XOR + POPC), bound by launch latency at SLAM sizes (2000 descriptors x 6 levels
// x 10 children = 120 k distances per frame).
#include <cstring>
#include <vector>

#include "../../include/orbfe_bow.h"
#include "orbfe_internal.h"

namespace orbfe {

__device__ __forceinline__ int bow_ham256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// One warp per descriptor; at every level the lanes take the children of the current node (10 in ORBvoc), the
// minimum of (distance << 16 | child position) over the warp is the reference's strict-< scan (first minimum wins).
__global__ void __launch_bounds__(256) bow_descend_kernel(const uint4 *__restrict__ node_desc, const int *__restrict__ child_ptr,
                                                          const int *__restrict__ children, const uint4 *__restrict__ desc,
                                                          int n, int nid_level, int *__restrict__ leaf_out,
                                                          int *__restrict__ node_out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const uint4 a0 = __ldg(&desc[2 * i]), a1 = __ldg(&desc[2 * i + 1]);
    int cur = 0, level = 0, nid = 0;
    int cb = __ldg(&child_ptr[0]), ce = __ldg(&child_ptr[1]);
    // do { ... } while (!isLeaf): the root of a non-empty vocabulary has children.  The level cap only guards against a
    // malformed table with a cycle (a DBoW2 tree is a handful of levels deep): the kernel must terminate.
    while (ce > cb && level < 64) {
        level++;
        uint32_t best = 0xFFFFFFFFu;
        for (int c0 = cb; c0 < ce; c0 += 32) {
            const int c = c0 + lane;
            uint32_t key = 0xFFFFFFFFu;
            if (c < ce) {
                const int id = __ldg(&children[c]);
                const int d = bow_ham256(a0, a1, __ldg(&node_desc[2 * id]), __ldg(&node_desc[2 * id + 1]));
                key = ((uint32_t)d << 16) | (uint32_t)min(c - cb, 0xFFFF);
            }
            best = min(best, __reduce_min_sync(0xffffffffu, key));
        }
        cur = __ldg(&children[cb + (int)(best & 0xFFFF)]);
        if (level == nid_level) nid = cur;
        cb = __ldg(&child_ptr[cur]);
        ce = __ldg(&child_ptr[cur + 1]);
    }
    if (lane == 0) {
        leaf_out[i] = cur;
        node_out[i] = nid;
    }
}

// One warp per map point.  Row i of the N x N distance matrix is histogrammed (257 bins) in shared memory and the
// (int)(0.5*(N-1))-th smallest entry read off the cumulative counts; the first row with the least median wins.
__global__ void __launch_bounds__(128) distinctive_kernel(const uint4 *__restrict__ desc, const int *__restrict__ group_ptr,
                                                          int ngroups, int *__restrict__ best_out) {
    __shared__ int hist[4][288];
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * 4 + wq;
    if (g >= ngroups) return;
    const int b = __ldg(&group_ptr[g]), N = __ldg(&group_ptr[g + 1]) - b;
    if (N <= 0) {
        if (lane == 0) best_out[g] = -1;
        return;
    }
    const int k = (N - 1) >> 1;  // vDists[0.5*(N-1)]
    int *h = hist[wq];
    int bestMedian = 0x7FFFFFFF, bestIdx = 0;
    for (int i = 0; i < N; i++) {
        for (int t = lane; t < 288; t += 32) h[t] = 0;
        __syncwarp();
        const uint4 a0 = __ldg(&desc[2 * (size_t)(b + i)]), a1 = __ldg(&desc[2 * (size_t)(b + i) + 1]);
        for (int j = lane; j < N; j += 32) {
            const int d = bow_ham256(a0, a1, __ldg(&desc[2 * (size_t)(b + j)]), __ldg(&desc[2 * (size_t)(b + j) + 1]));
            atomicAdd(&h[d], 1);
        }
        __syncwarp();
        // lane L owns bins 9L .. 9L+8 (288 >= 257 bins)
        int c[9], sum = 0;
#pragma unroll
        for (int t = 0; t < 9; t++) { c[t] = h[9 * lane + t]; sum += c[t]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int fl = __ffs(__ballot_sync(0xffffffffu, incl > k)) - 1;  // lane whose bins hold the k-th smallest
        int median = 0;
        if (lane == fl) {
            int run = incl - sum;
#pragma unroll
            for (int t = 0; t < 9; t++) {
                run += c[t];
                if (run > k) { median = 9 * lane + t; break; }
            }
        }
        median = __shfl_sync(0xffffffffu, median, fl);
        if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) best_out[g] = bestIdx;
}

// KeyFrameDatabase scoring: one thread per keyframe, two-pointer walk over the (ascending) word lists of the query and
// of the keyframe.  score accumulates |vi - wi| - |vi| - |wi| over the shared words in ascending word order, as
// L1Scoring::score does (ScoringObject.cpp:33-56; its lower_bound jumps visit the same shared words in the same order).
__global__ void __launch_bounds__(128) bow_db_score_kernel(int nq, const int *__restrict__ q_ids, const double *__restrict__ q_vals,
                                                           int nkf, const int *__restrict__ kf_ptr, const int *__restrict__ db_ids,
                                                           const double *__restrict__ db_vals, int *__restrict__ common_out,
                                                           int *__restrict__ first_out, double *__restrict__ score_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkf) return;
    int a = 0, b = __ldg(&kf_ptr[k]);
    const int be = __ldg(&kf_ptr[k + 1]);
    int common = 0, first = -1;
    double score = 0.0;
    while (a < nq && b < be) {
        const int ia = __ldg(&q_ids[a]), ib = __ldg(&db_ids[b]);
        if (ia == ib) {
            const double vi = __ldg(&q_vals[a]), wi = __ldg(&db_vals[b]);
            score = __dadd_rn(score, __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi)));
            if (common == 0) first = ia;
            common++;
            a++; b++;
        } else if (ia < ib) {
            a++;
        } else {
            b++;
        }
    }
    common_out[k] = common;
    first_out[k] = first;
    score_out[k] = -score / 2.0;
}

void launch_bow_db_score(int nq, const int *q_ids, const double *q_vals, int nkf, const int *kf_ptr, const int *db_ids,
                         const double *db_vals, int *common, int *first, double *score, cudaStream_t s) {
    if (nkf <= 0) return;
    bow_db_score_kernel<<<(nkf + 127) / 128, 128, 0, s>>>(nq, q_ids, q_vals, nkf, kf_ptr, db_ids, db_vals, common, first, score);
}

}  // namespace orbfe

using namespace orbfe;

#define BOW_TRY(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return set_error(ORBFE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct OrbfeVocabulary {
    int device = 0, nnodes = 0, depth = 0, weighting = 0, norm = 0;
    cudaStream_t stream = nullptr;
    uint8_t *d_desc = nullptr;
    int *d_child_ptr = nullptr, *d_children = nullptr;
    std::vector<int32_t> word_id;   // per node, host side (leaf -> word)
    std::vector<double> weight;
    // grow-only staging of the host-pointer entry point
    uint8_t *d_in = nullptr;
    int *d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
};

extern "C" void orbfe_vocabulary_destroy(OrbfeVocabulary *v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamDestroy(v->stream);
    cudaFree(v->d_desc); cudaFree(v->d_child_ptr); cudaFree(v->d_children); cudaFree(v->d_in); cudaFree(v->d_out);
    delete v;
}

extern "C" OrbfeVocabulary *orbfe_vocabulary_create(int device, int nnodes, int depth_L, const uint8_t *node_desc,
                                                    const int32_t *child_ptr, const int32_t *children, const int32_t *word_id,
                                                    const double *weight, int weighting, int norm) {
    if (nnodes < 1 || depth_L < 0 || !node_desc || !child_ptr || !word_id || !weight || weighting < 0 || weighting > 3 || norm < 0 ||
        norm > 2) {
        set_error(ORBFE_ERR_ARG, "bad vocabulary arguments");
        return nullptr;
    }
    const int nchild = child_ptr[nnodes];
    if (child_ptr[0] != 0 || nchild < 0 || (nchild > 0 && !children)) { set_error(ORBFE_ERR_ARG, "bad child_ptr"); return nullptr; }
    for (int i = 0; i < nnodes; i++)
        if (child_ptr[i + 1] < child_ptr[i]) { set_error(ORBFE_ERR_ARG, "child_ptr must be non-decreasing"); return nullptr; }
    for (int c = 0; c < nchild; c++)
        if (children[c] <= 0 || children[c] >= nnodes) { set_error(ORBFE_ERR_ARG, "child id out of range"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error(ORBFE_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { set_error(ORBFE_ERR_ARG, "device %d out of range", device); return nullptr; }
    OrbfeVocabulary *v = new OrbfeVocabulary();
    v->device = device; v->nnodes = nnodes; v->depth = depth_L; v->weighting = weighting; v->norm = norm;
    v->word_id.assign(word_id, word_id + nnodes);
    v->weight.assign(weight, weight + nnodes);
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc((void **)&v->d_desc, (size_t)nnodes * 32) == cudaSuccess &&
              cudaMalloc((void **)&v->d_child_ptr, sizeof(int) * ((size_t)nnodes + 1)) == cudaSuccess &&
              cudaMalloc((void **)&v->d_children, sizeof(int) * (size_t)std::max(nchild, 1)) == cudaSuccess &&
              cudaMemcpy(v->d_desc, node_desc, (size_t)nnodes * 32, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(v->d_child_ptr, child_ptr, sizeof(int) * ((size_t)nnodes + 1), cudaMemcpyHostToDevice) == cudaSuccess &&
              (nchild == 0 || cudaMemcpy(v->d_children, children, sizeof(int) * (size_t)nchild, cudaMemcpyHostToDevice) == cudaSuccess);
    if (!ok) {
        set_error(ORBFE_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        orbfe_vocabulary_destroy(v);
        return nullptr;
    }
    return v;
}

extern "C" int orbfe_bow_descend_device(OrbfeVocabulary *v, const uint8_t *d_desc, int n, int levelsup, int32_t *d_leaf_out,
                                        int32_t *d_node_out, void *stream) {
    if (!v || n < 0) return set_error(ORBFE_ERR_ARG, "bad arguments");
    if (n == 0) return ORBFE_OK;
    if (!d_desc || !d_leaf_out || !d_node_out) return set_error(ORBFE_ERR_ARG, "NULL argument");
    BOW_TRY(cudaSetDevice(v->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : v->stream;
    bow_descend_kernel<<<(n + 7) / 8, 256, 0, s>>>(reinterpret_cast<const uint4 *>(v->d_desc), v->d_child_ptr, v->d_children,
                                                   reinterpret_cast<const uint4 *>(d_desc), n, v->depth - levelsup, d_leaf_out,
                                                   d_node_out);
    BOW_TRY(cudaGetLastError());
    return ORBFE_OK;
}

extern "C" int orbfe_bow_descend(OrbfeVocabulary *v, const uint8_t *desc, int n, int levelsup, int32_t *leaf_out, int32_t *node_out) {
    if (!v || n < 0) return set_error(ORBFE_ERR_ARG, "bad arguments");
    if (n == 0) return ORBFE_OK;
    if (!desc || !leaf_out || !node_out) return set_error(ORBFE_ERR_ARG, "NULL argument");
    BOW_TRY(cudaSetDevice(v->device));
    if (v->in_cap < (size_t)n * 32) {
        cudaFree(v->d_in); v->d_in = nullptr; v->in_cap = 0;
        BOW_TRY(cudaMalloc((void **)&v->d_in, (size_t)n * 32 * 2));
        v->in_cap = (size_t)n * 32 * 2;
    }
    if (v->out_cap < (size_t)n * 2) {
        cudaFree(v->d_out); v->d_out = nullptr; v->out_cap = 0;
        BOW_TRY(cudaMalloc((void **)&v->d_out, sizeof(int) * (size_t)n * 4));
        v->out_cap = (size_t)n * 4;
    }
    BOW_TRY(cudaMemcpyAsync(v->d_in, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    int rc = orbfe_bow_descend_device(v, v->d_in, n, levelsup, v->d_out, v->d_out + n, v->stream);
    if (rc) return rc;
    BOW_TRY(cudaMemcpyAsync(leaf_out, v->d_out, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, v->stream));
    BOW_TRY(cudaMemcpyAsync(node_out, v->d_out + n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, v->stream));
    BOW_TRY(cudaStreamSynchronize(v->stream));
    return ORBFE_OK;
}

// accessors for host/bow_host.cpp (the struct layout stays private to this file)
namespace orbfe {
const int32_t *vocab_word_ids(const OrbfeVocabulary *v) { return v->word_id.data(); }
const double *vocab_weights(const OrbfeVocabulary *v) { return v->weight.data(); }
void vocab_modes(const OrbfeVocabulary *v, int *weighting, int *norm) { *weighting = v->weighting; *norm = v->norm; }

void launch_distinctive(const uint8_t *d_desc, const int *d_group_ptr, int ngroups, int *d_best, cudaStream_t s) {
    if (ngroups <= 0) return;
    distinctive_kernel<<<(ngroups + 3) / 4, 128, 0, s>>>(reinterpret_cast<const uint4 *>(d_desc), d_group_ptr, ngroups, d_best);
}
}  // namespace orbfe
