// mma_lab.cu — latency / throughput of the legacy warp-level MMA and dot-product instructions on sm_100a, the
// numbers the dequant-GEMV design depends on: IMMA.16832 (u8 x s8), HMMA.16816 (f16 -> f32), IDP.4A (dp4a).
// One CTA per SM, W warps per CTA; every warp runs `chains` independent dependent-chains of `iters` instructions.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>

__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KIND, int CH>
__global__ void k(int iters, long long* cyc, int* sink) {
    uint32_t a0 = threadIdx.x * 3 + 1, a1 = a0 * 5, a2 = a0 * 7, a3 = a0 * 11, b0 = a0 * 13, b1 = a0 * 17;
    int ci[CH][4];
    float cf[CH][4];
#pragma unroll
    for (int j = 0; j < CH; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) { ci[j][q] = j + q; cf[j][q] = (float)(j + q); }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < CH; j++) {
            if (KIND == 0) imma(ci[j], a0, a1, a2, a3, b0, b1);
            else if (KIND == 1) hmma(cf[j], a0, a1, a2, a3, b0, b1);
            else {
#pragma unroll
                for (int q = 0; q < 4; q++) ci[j][q] = __dp4a((int)a0, (int)b0, ci[j][q]);
            }
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int j = 0; j < CH; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc += ci[j][q] + (int)cf[j][q];
    if (acc == 0x7fffffff) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int KIND, int CH>
static void run(const char* name, int warps) {
    long long* cyc;
    int* sink;
    cudaMalloc(&cyc, 8);
    cudaMalloc(&sink, 4);
    const int iters = 2000;
    k<KIND, CH><<<148, warps * 32>>>(iters, cyc, sink);
    k<KIND, CH><<<148, warps * 32>>>(iters, cyc, sink);
    long long h = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / iters;                  // cycles per round of CH instructions (x4 for dp4a)
    const int n = CH * (KIND == 2 ? 4 : 1);
    printf("%-6s warps/SM %2d chains %d: %7.1f clk per round, %6.2f clk per instr per warp, %6.3f warp-instr/clk/SM\n", name, warps, CH, per,
           per / n, (double)n * warps / per);
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    for (int w : {1, 4, 8, 16}) {
        run<0, 1>("IMMA", w); run<0, 4>("IMMA", w); run<0, 8>("IMMA", w);
        run<1, 1>("HMMA", w); run<1, 4>("HMMA", w); run<1, 8>("HMMA", w);
        run<2, 1>("DP4A", w); run<2, 4>("DP4A", w);
    }
    return 0;
}
