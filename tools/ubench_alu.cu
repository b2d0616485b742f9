// ubench_alu.cu -- issue-rate micro-benchmark for the integer ops the FAST kernel is built from (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_alu tools/ubench_alu.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
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
    double ops = (double)148 * 4 * 512 * ITERS * 8;  // thread-ops of the op under test
    printf("%-22s %8.3f ms  %7.2f Tthread-op/s  %6.2f thread-ops/clk/SM (at 1.965 GHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("vimin3_s16x2+iadd"); run<1>("vimin3_s16x2"); run<2>("vimin3_s32"); run<3>("vabsdiffu4"); run<4>("prmt");
    run<5>("lop3"); run<6>("iadd3"); run<7>("imad"); run<8>("vimax_s16x2"); run<9>("popc+iadd"); run<10>("shf");
    return 0;
}
