// tma_test.cu -- standalone check of the 3-D u8 TMA tile load used by fast_nms_tma_kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tma_test tools/tma_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BOXW, int BOXH>
__global__ void k(const CUtensorMap *gmap, const __grid_constant__ CUtensorMap pmap, int use_param, int cx, int cy, int cz, unsigned char *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BOXW * BOXH) : "memory");
        const void *tm = use_param ? (const void *)&pmap : (const void *)gmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(sm)), "l"(tm), "r"(smem_u32(&bar)), "r"(cx), "r"(cy), "r"(cz) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < BOXW * BOXH; i += blockDim.x) out[i] = sm[i];
}
template <int BOXW, int BOXH>
int run(int W, int H, int pitch, int B, int use_param, int cx = 8) {
    std::vector<unsigned char> img((size_t)pitch * H * B);
    for (size_t i = 0; i < img.size(); i++) img[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d_img, *d_out;
    cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, BOXW * BOXH);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
    cuuint32_t box[3] = {BOXW, BOXH, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    CUtensorMap *d_m; cudaMalloc(&d_m, sizeof(m)); cudaMemcpy(d_m, &m, sizeof(m), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<BOXW, BOXH>, cudaFuncAttributeMaxDynamicSharedMemorySize, BOXW * BOXH + 128);
    int cy = 12, cz = B - 1;
    k<BOXW, BOXH><<<1, 128, BOXW * BOXH + 128>>>(d_m, m, use_param, cx, cy, cz, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box %dx%d W=%d H=%d pitch=%d B=%d param=%d cx=%d: %s", BOXW, BOXH, W, H, pitch, B, use_param, cx, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<unsigned char> out(BOXW * BOXH);
        cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < BOXH; y++) for (int x = 0; x < BOXW; x++) {
            int gx = cx + x, gy = cy + y;
            unsigned char ref = (gx < W && gy < H) ? img[(size_t)cz * pitch * H + (size_t)gy * pitch + gx] : 0;
            bad += out[y * BOXW + x] != ref;
        }
        printf("  mismatches %d", bad);
    }
    printf("\n");
    return e != cudaSuccess;
}
int main(int argc, char **argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 5) return run<144, 70>(640, 480, 640, 1, 1, 16);
    if (which == 6) return run<144, 70>(640, 480, 640, 2, 0, 0);
    if (which == 7) return run<272, 38>(1920, 1080, 1920, 2, 0, 240);
    if (which == 8) return run<144, 70>(640, 480, 640, 2, 0, 624);
    if (which == 0) return run<144, 70>(640, 480, 640, 1, 1);
    if (which == 1) return run<144, 70>(640, 480, 640, 1, 0);
    if (which == 2) return run<128, 70>(640, 480, 640, 2, 1);
    if (which == 3) return run<144, 64>(640, 480, 640, 2, 1);
    if (which == 4) return run<144, 70>(1920, 1080, 1920, 4, 0);
    return 0;
}
