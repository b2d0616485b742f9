// ubench_alu.cu -- issue-rate micro-benchmark for the integer ops the FAST kernel is built from (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_alu tools/ubench_alu.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) { unsigned d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hmin2u(unsigned a, unsigned b) { unsigned d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hfma2u(unsigned a, unsigned b, unsigned c) { unsigned d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned hfma2relu(unsigned a, unsigned b, unsigned c) { unsigned d; asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__global__ void denorm_check(unsigned *o) {
    // u16 values 0..255 are fp16 subnormals: max/min must order them like integers and return them unflushed
    unsigned bad = 0;
    for (unsigned x = threadIdx.x; x < 65536; x += blockDim.x) {
        const unsigned a = (x & 0xFF) | ((x >> 8) << 16), b = ((x * 7 + 3) & 0xFF) | (((x * 13 + 5) & 0xFF) << 16);
        const unsigned mx = hmax2u(a, b), mn = hmin2u(a, b);
        if (mx != __vmaxu2(a, b) || mn != __vminu2(a, b)) bad++;
    }
    atomicAdd(o, bad);
}
template <int OP>
__global__ void k(unsigned *out, unsigned seed) {
    unsigned a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    unsigned b = seed ^ 0x00ff00ffu, c = seed + 0x01000100u;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __vimin3_s16x2(a[i], b, c) + 1;            // VIMNMX3.S16x2 (+IADD)
            if (OP == 1) a[i] = __vimin3_s16x2(a[i], b, c);                // pure
            if (OP == 2) a[i] = (unsigned)__vimin3_s32((int)a[i], (int)b, (int)c);
            if (OP == 3) a[i] = __vabsdiffu4(a[i], b);
            if (OP == 4) a[i] = __byte_perm(a[i], b, 0x3254);
            if (OP == 5) a[i] = (a[i] & b) | c;                            // LOP3
            if (OP == 6) a[i] = a[i] + b + c;                              // IADD3
            if (OP == 7) a[i] = a[i] * b + c;                              // IMAD
            if (OP == 8) a[i] = __vmaxs2(a[i], b);                    // 2-input
            if (OP == 9) a[i] = __popc(a[i]) + b;                          // POPC
            if (OP == 10) a[i] = __funnelshift_r(a[i], b, 8);              // SHF
            if (OP == 11) a[i] = hmax2u(a[i], b);                          // HMNMX2 on u16x2 bit patterns (< 0x7C00: ordered like fp16)
            if (OP == 12) a[i] = (i & 1) ? hmax2u(a[i], b) : __vimax3_u16x2(a[i], b, c);   // 4 + 4: different pipes?
            if (OP == 13) a[i] = (i & 1) ? a[i] * b + c : __vimax3_u16x2(a[i], b, c);     // VIMNMX3 + IMAD (ALU + FMA pipe)
            if (OP == 14) a[i] = __float_as_uint(fmax3(__uint_as_float(a[i]), __uint_as_float(b), __uint_as_float(c)));
            if (OP == 15) a[i] = (i % 3 == 2) ? hmax2u(a[i], b) : __vimax3_u16x2(a[i], b, c);  // 5-6 VIMNMX3 + 2-3 HMNMX2
            if (OP == 17) a[i] = hfma2u(a[i], 0xBC00BC00u, b);                          // HFMA2 (FMA pipe), fp16 subnormal operands
            if (OP == 18) a[i] = (i & 1) ? hfma2relu(a[i], 0xBC00BC00u, b) : __vimax3_u16x2(a[i], b, c);   // 4 VIMNMX3 + 4 HFMA2
            if (OP == 19) a[i] = (i % 3 == 2) ? hfma2relu(a[i], 0xBC00BC00u, b) : __vimax3_u16x2(a[i], b, c);   // ~5.3 VIMNMX3 + 2.7 HFMA2
            if (OP == 20) a[i] = (i & 1) ? hfma2u(hfma2relu(a[i], 0xBC00BC00u, b), 0x3C003C00u, c) : __vimax3_u16x2(a[i], b, c);   // 4 VIMNMX3 + 8 HFMA2
            if (OP == 21) a[i] = (i & 3) ? __vimax3_u16x2(a[i], b, c) : hfma2relu(a[i], 0xBC00BC00u, b);   // 6 VIMNMX3 + 2 HFMA2
            if (OP == 16) a[i] = (i & 1) ? hmin2u(hmax2u(a[i], b), c) : __vimax3_u16x2(a[i], b, c);   // 4 VIMNMX3 + 8 HMNMX2
        }
        b += 0x00010001u;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name) {
    unsigned *out;
    cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 4, 512>>>(out, 12345u);
    cudaEventRecord(e0);
    k<OP><<<148 * 4, 512>>>(out, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)148 * 4 * 512 * ITERS * (OP == 20 ? 12 : 8);  // thread-ops of the op under test (all pipes)
    printf("%-22s %8.3f ms  %7.2f Tthread-op/s  %6.2f thread-ops/clk/SM (at 1.965 GHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("vimin3_s16x2+iadd"); run<1>("vimin3_s16x2"); run<2>("vimin3_s32"); run<3>("vabsdiffu4"); run<4>("prmt");
    {
        unsigned *d, hbad = 123;
        cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
        denorm_check<<<1, 256>>>(d);
        cudaMemcpy(&hbad, d, 4, cudaMemcpyDeviceToHost);
        printf("HMNMX2 on u16x2 subnormal bit patterns: %u mismatches vs vmaxu2/vminu2\n", hbad);
    }
    run<11>("hmnmx2"); run<12>("4 vimnmx3 + 4 hmnmx2"); run<13>("4 vimnmx3 + 4 imad"); run<14>("fmnmx3"); run<15>("vimnmx3:hmnmx2 ~2:1");
    run<16>("4 vimnmx3 + 8 hmnmx2");
    run<17>("hfma2 (subnormal)"); run<18>("4 vimnmx3 + 4 hfma2"); run<19>("~5.3 vimnmx3 + 2.7 hfma2"); run<20>("4 vimnmx3 + 8 hfma2"); run<21>("6 vimnmx3 + 2 hfma2");
    run<5>("lop3"); run<6>("iadd3"); run<7>("imad"); run<8>("vimax_s16x2"); run<9>("popc+iadd"); run<10>("shf");
    return 0;
}
