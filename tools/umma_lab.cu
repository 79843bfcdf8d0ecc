// umma_lab.cu — first tcgen05 / TMEM dequant-GEMM for the prefill / batched-decode contraction (north_star (2)):
//   Y[t][j] = sum_k deq(W)[j][k] * X[t][k],  W = Q4_K rows in the untouched GGUF layout, X f32, Y f32.
// One CTA per (128 weight rows, TN tokens) tile.  Per K step of 64 elements (a quarter of a Q4_K super-block: the low
// and high nibbles of 32 qs bytes) every thread dequantises ITS weight row in registers (reference arithmetic
// d*sc*q - dmin*m in f32, dequant.rs:205-256), rounds to fp16 and writes the 128-byte row into the 128B-swizzled K-major
// shared-memory tile the UMMA descriptor describes; the activations go in the same way; one elected thread issues
// 4 x tcgen05.mma.cta_group::1.kind::f16 (M = 128 weight rows, N = TN tokens, K = 16), accumulators in TMEM;
// tcgen05.commit -> mbarrier frees the tiles; the epilogue reads TMEM with tcgen05.ld.32x32b and stores Y coalesced.
// Stand-alone: checks against a double-precision dequant-then-dot on the host and reports TFLOP/s.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int BM = 128;   // weight rows per CTA (UMMA M)
constexpr int BK = 64;    // K elements per step (128 bytes of fp16: one swizzle atom row)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: start >> 4 in [0,14), LBO >> 4 in
// [16,30), SBO >> 4 in [32,46), version = 1 in [46,48), layout type 2 = SWIZZLE_128B in [61,64)).  Rows are 128 bytes,
// 8-row groups are 1024 bytes apart (SBO); LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) break;   // never hang the box
    }
}

__device__ __forceinline__ float h2f(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)(h & 0xffffu))); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a 128B-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

template <int TN>
__global__ void __launch_bounds__(128) dqgemm_q4k(const uint8_t* __restrict__ W, long long row_bytes, int n_rows, int K, const float* __restrict__ X,
                                                  int T, float* __restrict__ Y) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;
    uint8_t* sA = smem;                    // [128 rows][128 B] swizzled
    uint8_t* sB = smem + BM * 128;         // [TN rows][128 B] swizzled
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row0 = blockIdx.x * BM, tok0 = blockIdx.y * TN;
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TN < 32 ? 32 : TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(BM, TN);
    const int my_row = min(row0 + tid, n_rows - 1);
    const uint8_t* wrow = W + (long long)my_row * row_bytes;
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- A: this thread's weight row, 64 elements of block k0/256, group gp = (k0 % 256) / 64 ----
        {
            const uint8_t* blk = wrow + (long long)(k0 >> 8) * 144;
            const int gp = (k0 & 255) >> 6;
            const uint32_t dd = *reinterpret_cast<const uint32_t*>(blk);
            const float d = h2f(dd), dmin = h2f(dd >> 16);
            const uint8_t* sc = blk + 4;
            int s1, m1, s2, m2;   // get_scale_min_k4 (dequant.rs:213-225) for sub-blocks 2gp, 2gp+1
            {
                const int j = 2 * gp;
                if (j < 4) { s1 = sc[j] & 63; m1 = sc[j + 4] & 63; } else { s1 = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m1 = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
                const int j2 = j + 1;
                if (j2 < 4) { s2 = sc[j2] & 63; m2 = sc[j2 + 4] & 63; } else { s2 = (sc[j2 + 4] & 0xF) | ((sc[j2 - 4] >> 6) << 4); m2 = (sc[j2 + 4] >> 4) | ((sc[j2] >> 6) << 4); }
            }
            const float d1 = d * (float)s1, mm1 = dmin * (float)m1, d2 = d * (float)s2, mm2 = dmin * (float)m2;
            const uint4* q4 = reinterpret_cast<const uint4*>(blk + 16 + 32 * gp);
            const uint4 qa = q4[0], qb = q4[1];
            const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            // elements 0..31 = low nibbles of bytes 0..31, elements 32..63 = high nibbles
#pragma unroll
            for (int c = 0; c < 4; c++) {   // chunk c: elements 8c..8c+7 (low), chunk 4+c: elements 32+8c.. (high)
                uint32_t lo[4], hi[4];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t w = qw[2 * c + h];
                    const float l0 = __fsub_rn(__fmul_rn(d1, (float)(w & 15)), mm1), l1 = __fsub_rn(__fmul_rn(d1, (float)((w >> 8) & 15)), mm1);
                    const float l2 = __fsub_rn(__fmul_rn(d1, (float)((w >> 16) & 15)), mm1), l3 = __fsub_rn(__fmul_rn(d1, (float)((w >> 24) & 15)), mm1);
                    const float h0 = __fsub_rn(__fmul_rn(d2, (float)((w >> 4) & 15)), mm2), h1 = __fsub_rn(__fmul_rn(d2, (float)((w >> 12) & 15)), mm2);
                    const float h2 = __fsub_rn(__fmul_rn(d2, (float)((w >> 20) & 15)), mm2), h3 = __fsub_rn(__fmul_rn(d2, (float)((w >> 28) & 15)), mm2);
                    lo[2 * h] = pack_h2(l0, l1); lo[2 * h + 1] = pack_h2(l2, l3);
                    hi[2 * h] = pack_h2(h0, h1); hi[2 * h + 1] = pack_h2(h2, h3);
                }
                *reinterpret_cast<uint4*>(sA + sw128(tid, c)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(sA + sw128(tid, 4 + c)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            }
        }
        // ---- B: TN tokens x 64 elements of X, f32 -> fp16; thread handles 16-byte chunks (8 elements) ----
        for (int i = tid; i < TN * 8; i += 128) {
            const int r = i >> 3, c = i & 7;
            const int tk = tok0 + r;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (tk < T) {
                const float4 a = *reinterpret_cast<const float4*>(X + (long long)tk * K + k0 + 8 * c);
                const float4 b = *reinterpret_cast<const float4*>(X + (long long)tk * K + k0 + 8 * c + 4);
                v = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
            }
            *reinterpret_cast<uint4*>(sB + sw128(r, c)) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = umma_desc(smem_u32(sA)), db = umma_desc(smem_u32(sB));
#pragma unroll
            for (int kk = 0; kk < BK / 16; kk++)   // 16 fp16 = 32 bytes further along the swizzled row: start address + 2
                umma_f16(tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc, (k0 > 0 || kk > 0) ? 1u : 0u);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);   // the MMAs have read the tiles (and, after the last step, written the accumulator)
        phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: warp w reads TMEM lanes 32w..32w+31 (weight rows), 32 token columns at a time ----
    const int j = row0 + tid;
#pragma unroll 1
    for (int n0 = 0; n0 < TN; n0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int n = 0; n < 32; n++) {
            const int tk = tok0 + n0 + n;
            if (tk < T && j < n_rows) Y[(long long)tk * n_rows + j] = __uint_as_float(v[n]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TN < 32 ? 32 : TN) : "memory");
}

// ---------------------------------------------------------------- host
static float h_h2f(uint16_t h) {
    const uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
    float v;
    if (e == 0) v = ldexpf((float)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexpf((float)(m | 1024), (int)e - 25);
    return s ? -v : v;
}
static uint16_t h_f2h(float f) {   // round to nearest even, normal range only (test scales)
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t s = (x >> 16) & 0x8000;
    int e = (int)((x >> 23) & 255) - 127 + 15;
    uint32_t m = x & 0x7fffff;
    if (e <= 0) return (uint16_t)s;
    uint32_t r = (m >> 13) + (((m & 0x1fff) > 0x1000 || ((m & 0x1fff) == 0x1000 && ((m >> 13) & 1))) ? 1 : 0);
    if (r == 1024) { r = 0; e++; }
    return (uint16_t)(s | (e << 10) | r);
}
static void deq_q4k_row(const uint8_t* row, int K, std::vector<float>& out) {
    out.resize(K);
    for (int b = 0; b < K / 256; b++) {
        const uint8_t* blk = row + (size_t)b * 144;
        uint16_t dh, mh;
        memcpy(&dh, blk, 2);
        memcpy(&mh, blk + 2, 2);
        const float d = h_h2f(dh), dmin = h_h2f(mh);
        const uint8_t* sc = blk + 4;
        const uint8_t* qs = blk + 16;
        for (int gp = 0; gp < 4; gp++) {
            int s[2], m[2];
            for (int h = 0; h < 2; h++) {
                const int j = 2 * gp + h;
                if (j < 4) { s[h] = sc[j] & 63; m[h] = sc[j + 4] & 63; } else { s[h] = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m[h] = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
            }
            for (int l = 0; l < 32; l++) {
                out[b * 256 + gp * 64 + l] = d * (float)s[0] * (float)(qs[32 * gp + l] & 15) - dmin * (float)m[0];
                out[b * 256 + gp * 64 + 32 + l] = d * (float)s[1] * (float)(qs[32 * gp + l] >> 4) - dmin * (float)m[1];
            }
        }
    }
}

template <int TN>
static int run(int n_rows, int K, int T, int iters) {
    const long long row_bytes = (long long)(K / 256) * 144;
    std::vector<uint8_t> hW((size_t)n_rows * row_bytes);
    srand(1234);
    for (auto& b : hW) b = (uint8_t)(rand() & 255);
    for (int r = 0; r < n_rows; r++)
        for (int b = 0; b < K / 256; b++) {
            const uint16_t d = h_f2h(0.002f + 0.001f * (float)((r + b) % 7)), m = h_f2h(0.01f + 0.002f * (float)((r * 3 + b) % 5));
            memcpy(&hW[(size_t)r * row_bytes + (size_t)b * 144], &d, 2);
            memcpy(&hW[(size_t)r * row_bytes + (size_t)b * 144 + 2], &m, 2);
        }
    std::vector<float> hX((size_t)T * K);
    for (auto& v : hX) v = (float)((rand() % 2001) - 1000) / 1000.0f;
    uint8_t* dW;
    float *dX, *dY;
    CK(cudaMalloc(&dW, hW.size() + 256));
    CK(cudaMalloc(&dX, hX.size() * 4));
    CK(cudaMalloc(&dY, (size_t)T * n_rows * 4));
    CK(cudaMemcpy(dW, hW.data(), hW.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dX, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dY, 0xFF, (size_t)T * n_rows * 4));
    const size_t smem = (size_t)(BM + TN) * 128 + 1024;
    CK(cudaFuncSetAttribute(dqgemm_q4k<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n_rows + BM - 1) / BM, (T + TN - 1) / TN);
    dqgemm_q4k<TN><<<grid, 128, smem>>>(dW, row_bytes, n_rows, K, dX, T, dY);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("TN %d rows %d K %d T %d: %s\n", TN, n_rows, K, T, cudaGetErrorString(e)); return 2; }
    std::vector<float> hY((size_t)T * n_rows);
    CK(cudaMemcpy(hY.data(), dY, hY.size() * 4, cudaMemcpyDeviceToHost));
    // check a sample of rows against double precision
    double max_err = 0, max_ref = 0;
    std::vector<float> wr;
    for (int r = 0; r < n_rows; r += (n_rows > 512 ? 37 : 1)) {
        deq_q4k_row(&hW[(size_t)r * row_bytes], K, wr);
        for (int t = 0; t < T; t += (T > 64 ? 5 : 1)) {
            double acc = 0;
            for (int k = 0; k < K; k++) acc += (double)wr[k] * (double)hX[(size_t)t * K + k];
            max_ref = fmax(max_ref, fabs(acc));
            max_err = fmax(max_err, fabs(acc - (double)hY[(size_t)t * n_rows + r]));
        }
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) dqgemm_q4k<TN><<<grid, 128, smem>>>(dW, row_bytes, n_rows, K, dX, T, dY);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * n_rows * (double)K * T * iters;
    printf("TN %3d rows %6d K %5d T %4d: max|err| %.3e / max|ref| %.3e = %.2e   %.3f ms  %.1f TFLOP/s  (weights %.1f GB/s)\n", TN, n_rows, K, T, max_err,
           max_ref, max_err / max_ref, ms / iters, fl / (ms * 1e-3) / 1e12, (double)hW.size() * grid.y * iters / (ms * 1e-3) / 1e9);
    cudaFree(dW);
    cudaFree(dX);
    cudaFree(dY);
    return 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) return run<32>(128, 256, 32, 1);          // smallest: one CTA, one super-block
    if (which == 1) return run<32>(256, 1024, 32, 1);
    if (which == 2) return run<128>(4096, 4096, 128, 5);
    if (which == 3) return run<256>(4096, 4096, 512, 5);
    if (which == 4) return run<32>(14336, 4096, 32, 5);      // batch-32 decode shape
    if (which == 5) return run<256>(14336, 4096, 2048, 3);   // 2K prefill, gate projection
    return 0;
}
