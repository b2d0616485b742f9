// match_kernels.cu -- 256-bit Hamming kernels behind ORBmatcher::SearchBy* (reference src/ORBmatcher.cc).
//
// The per-pair primitive is ORBmatcher::DescriptorDistance (ORBmatcher.cc:1794-1810): popcount of the XOR
// of two 32-byte descriptors.  On the device that is 8 x (XOR + POPC) on 32-bit words.  These kernels are
// bound by the integer/POPC issue rate (all-pairs sweep) or by launch latency (windowed CSR lists), never
// by HBM; no tensor cores (the north star forbids reshaping Hamming into a GEMM).
#include "orbfe_internal.h"

namespace orbfe {

__device__ __forceinline__ int ham256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// One warp per query row of the CSR candidate structure; lanes stride over that row's candidates.
__global__ void __launch_bounds__(256) hamming_csr_kernel(const uint4 *__restrict__ q, const uint4 *__restrict__ t,
                                                          const int32_t *__restrict__ row_ptr,
                                                          const int32_t *__restrict__ cols, int nq,
                                                          uint16_t *__restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nq) return;
    const int lane = threadIdx.x & 31;
    const int beg = __ldg(&row_ptr[row]), end = __ldg(&row_ptr[row + 1]);
    if (beg >= end) return;
    const uint4 a0 = __ldg(&q[2 * row]), a1 = __ldg(&q[2 * row + 1]);
    for (int k = beg + lane; k < end; k += 32) {
        const int j = __ldg(&cols[k]);
        const uint4 b0 = __ldg(&t[2 * j]), b1 = __ldg(&t[2 * j + 1]);
        out[k] = (uint16_t)ham256(a0, a1, b0, b1);
    }
}

void launch_hamming_csr(const uint8_t *q, const uint8_t *t, const int32_t *row_ptr, const int32_t *cols, int nq,
                        int npairs, uint16_t *out, cudaStream_t s) {
    (void)npairs;
    if (nq <= 0) return;
    hamming_csr_kernel<<<(nq + 7) / 8, 256, 0, s>>>(reinterpret_cast<const uint4 *>(q), reinterpret_cast<const uint4 *>(t),
                                                    row_ptr, cols, nq, out);
}

// Dense nq x nt matrix: blockIdx.y = query, threads stride over train descriptors.
__global__ void __launch_bounds__(256) hamming_dense_kernel(const uint4 *__restrict__ q, int nq,
                                                            const uint4 *__restrict__ t, int nt,
                                                            uint16_t *__restrict__ out) {
    const int i = blockIdx.y;
    const uint4 a0 = __ldg(&q[2 * i]), a1 = __ldg(&q[2 * i + 1]);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nt; j += gridDim.x * blockDim.x) {
        const uint4 b0 = __ldg(&t[2 * j]), b1 = __ldg(&t[2 * j + 1]);
        out[(size_t)i * nt + j] = (uint16_t)ham256(a0, a1, b0, b1);
    }
}

void launch_hamming_dense(const uint8_t *q, int nq, const uint8_t *t, int nt, uint16_t *out, cudaStream_t s) {
    if (nq <= 0 || nt <= 0) return;
    dim3 grid(min((nt + 255) / 256, 64), nq);
    hamming_dense_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4 *>(q), nq, reinterpret_cast<const uint4 *>(t), nt, out);
}

// Best / second-best sweep: blockIdx.x = database group (keyframe), blockIdx.y = block of 256 queries.
// Each thread owns one query in registers; the group's descriptors stream through shared memory in
// chunks and are read by all threads at the same address (broadcast).  Update rule = the strict-<
// best/second loop of ORBmatcher.cc:456-466, in ascending database order (first minimum wins).
#define KNN_CHUNK 128
__global__ void __launch_bounds__(256) knn2_groups_kernel(const uint4 *__restrict__ q, int nq,
                                                          const uint4 *__restrict__ db, int group_size,
                                                          uint16_t *__restrict__ best, int32_t *__restrict__ best_idx,
                                                          uint16_t *__restrict__ second) {
    __shared__ uint4 chunk[KNN_CHUNK * 2];
    const int g = blockIdx.x;
    const int qi = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = qi < nq;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (active) { a0 = __ldg(&q[2 * qi]); a1 = __ldg(&q[2 * qi + 1]); }
    int b1 = 0x7fffffff, b2 = 0x7fffffff, bi = -1;
    const uint4 *__restrict__ gdb = db + (size_t)g * group_size * 2;
    for (int c0 = 0; c0 < group_size; c0 += KNN_CHUNK) {
        const int nc = min(KNN_CHUNK, group_size - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * 2; i += blockDim.x) chunk[i] = __ldg(&gdb[(size_t)c0 * 2 + i]);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int j = 0; j < nc; j++) {
                const int d = ham256(a0, a1, chunk[2 * j], chunk[2 * j + 1]);
                if (d < b1) { b2 = b1; b1 = d; bi = c0 + j; }
                else if (d < b2) b2 = d;
            }
        }
    }
    if (active) {
        const size_t o = (size_t)g * nq + qi;
        best[o] = (uint16_t)min(b1, 0xFFFF);
        best_idx[o] = bi;
        second[o] = (uint16_t)min(b2, 0xFFFF);
    }
}

void launch_knn2_groups(const uint8_t *q, int nq, const uint8_t *db, int ngroups, int group_size,
                        uint16_t *best, int32_t *best_idx, uint16_t *second, cudaStream_t s) {
    if (nq <= 0 || ngroups <= 0) return;
    dim3 grid(ngroups, (nq + 255) / 256);
    knn2_groups_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4 *>(q), nq, reinterpret_cast<const uint4 *>(db),
                                            group_size, best, best_idx, second);
}

}  // namespace orbfe
