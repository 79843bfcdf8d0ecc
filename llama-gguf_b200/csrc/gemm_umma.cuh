// gemm_umma.cuh — tcgen05 / TMEM dequant-GEMM: the one dense contraction of prefill and batched decode.
//
//   Y[t][j] (+)= sum_k deq(W)[j][k] * X[t][k] (+ bias[j]),   W = N rows of GGUF blocks (Q4_K / Q5_K / Q6_K / Q8_0), X f32.
//
// The reference runs a prompt token by token through its vec_mat_q kernels (src/model/llama.rs:327-345,
// src/backend/cuda/gpu_only.rs:776-790: T GEMVs per weight matrix, every one of them re-reading the matrix); this is
// the same arithmetic as T columns of one GEMM: the weights are read once per 256-token tile.
//
// One CTA (128 threads) per (128 weight rows, TN tokens) output tile.  Per K step of 64 elements (128 bytes of fp16 =
// one row of a 128B-swizzle atom):
//   * (small T, TMA-fed: one cp.async.bulk.tensor.2d box of 128 rows x one 256-element GGUF block per row lands in a
//     two-stage raw buffer per 4 K steps, and the threads read their row from shared memory; large T: straight from global)
//   * thread r dequantises 64 elements of ITS weight row in registers with the reference arithmetic in f32
//     (dequant.rs:205-356: d*sc*q - dmin*m, (d*sc)*q for Q6_K, q*d for Q8_0), rounds to fp16 and writes the row's eight
//     16-byte chunks into the K-major SWIZZLE_128B shared-memory tile (chunk c of row r at chunk c ^ (r & 7));
//   * the activations (fp16 in HBM, rounded once by their producer) are copied into a second tile the same way;
//   * fence.proxy.async, then ONE elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M = 128 weight rows,
//     N = TN tokens, K = 16 each; SASS UTCHMMA) on the two shared-memory descriptors; the f32 accumulator lives in TMEM
//     (TN columns x 128 lanes); tcgen05.commit -> mbarrier tells the CTA that the tiles may be overwritten;
//   * epilogue: warp w reads TMEM lanes 32w..32w+31 with tcgen05.ld.32x32b.x32 (SASS LDTM), thread = weight row, 32
//     tokens per load, and stores Y so that a warp writes 32 consecutive floats of one token's row.
// fp16 operands, f32 accumulation: measured 3e-4 of the largest output against a double-precision dequant-then-dot
// (tools/umma_lab.cu); the exact token-by-token path stays the default where bit-level greedy parity matters.
// Two stages: the MMAs of step k overlap the dequant of step k + 1; the dequant (CUDA cores), not the tensor pipe, is the
// limit.
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

constexpr int kUmmaM = 128;   // weight rows per CTA
constexpr int kUmmaK = 64;    // K elements per step
constexpr int kUmmaRawStage = kUmmaM * 272;   // raw (quantised) tile of 128 rows x one 256-element block, widest type

__device__ __forceinline__ uint32_t umma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start >> 4 [0,14), LBO >> 4 [16,30) (unused for swizzled K-major), SBO >> 4 [32,46) = 1024 bytes between 8-row groups,
// version 1 [46,48), layout type 2 = SWIZZLE_128B [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = F16 (0), K-major both, N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool umma_mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: never hang the box
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!ok && clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}
__device__ __forceinline__ uint32_t umma_pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a 128B-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t umma_sw128(int r, int c) {
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}
// weight bytes come either straight from global memory (read-only path) or from the raw tile a TMA load left in shared memory
template <bool SM, class T>
__device__ __forceinline__ T umma_ldw(const T* p) {
    if constexpr (SM) return *p; else return __ldg(p);
}
template <bool SM>
__device__ __forceinline__ uint32_t umma_ld16(const uint8_t* p) { return (uint32_t)umma_ldw<SM>(reinterpret_cast<const unsigned short*>(p)); }

// Eight consecutive elements k0 + 8c .. k0 + 8c + 7 (k0 % 64 == 0) of the weight row `row`, as four packed half2.
template <bool SM>
__device__ __forceinline__ uint4 umma_deq8(int type, const uint8_t* row, int k0, int c) {
    float v[8];
    if (type == T_Q4_K || type == T_Q5_K) {   // dequant.rs:205-315
        const bool q5 = type == T_Q5_K;
        const uint8_t* blk = row + (long long)(k0 >> 8) * (q5 ? 176 : 144);
        const int gp = (k0 & 255) >> 6, hi = c >> 2;        // sub-block 2gp (low nibbles) or 2gp + 1 (high nibbles)
        const uint32_t dd = umma_ldw<SM>(reinterpret_cast<const uint32_t*>(blk));
        const float d = half_bits_to_float(dd), dmin = half_bits_to_float(dd >> 16);
        int sc, mn;
        scale_min_k4(blk + 4, 2 * gp + hi, sc, mn);
        const float d1 = __fmul_rn(d, (float)sc), m1 = __fmul_rn(dmin, (float)mn);
        const uint8_t* qs = blk + (q5 ? 48 : 16) + 32 * gp + 8 * (c & 3);
        const uint2 w = umma_ldw<SM>(reinterpret_cast<const uint2*>(qs));
        uint2 h = make_uint2(0u, 0u);
        if (q5) h = umma_ldw<SM>(reinterpret_cast<const uint2*>(blk + 16 + 8 * (c & 3)));
        const int sh = 4 * hi, hb = 2 * gp + hi;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t ww = i < 4 ? w.x : w.y, hh = i < 4 ? h.x : h.y;
            int q = (int)((ww >> (8 * (i & 3) + sh)) & 15u);
            if (q5) q += (int)((hh >> (8 * (i & 3) + hb)) & 1u) << 4;
            v[i] = __fsub_rn(__fmul_rn(d1, (float)q), m1);
        }
    } else if (type == T_Q6_K) {   // dequant.rs:321-356; blocks are only 2-byte aligned
        const uint8_t* blk = row + (long long)(k0 >> 8) * 210;
        const int ks = (k0 & 255) >> 6, n = ks >> 1, qq = 2 * (ks & 1) + (c >> 2), l0 = 8 * (c & 3);
        const float d = half_bits_to_float(umma_ld16<SM>(blk + 208));
        const int sc = (int)(signed char)umma_ldw<SM>(blk + 192 + 8 * n + (l0 >> 4) + 2 * qq);
        const float ds = __fmul_rn(d, (float)sc);
        const uint8_t* ql = blk + 64 * n + 32 * (qq & 1) + l0;
        const uint8_t* qh = blk + 128 + 32 * n + l0;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const uint32_t a = umma_ld16<SM>(ql + i), b = umma_ld16<SM>(qh + i);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t lo = (a >> (8 * j + 4 * (qq >> 1))) & 15u, hb = (b >> (8 * j + 2 * qq)) & 3u;
                v[i + j] = __fmul_rn(ds, (float)((int)(lo | (hb << 4)) - 32));
            }
        }
    } else {   // Q8_0, dequant.rs:103-109
        const uint8_t* blk = row + (long long)((k0 >> 5) + (c >> 2)) * 34;
        const float d = half_bits_to_float(umma_ld16<SM>(blk));
        const uint8_t* q = blk + 2 + 8 * (c & 3);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const uint32_t a = umma_ld16<SM>(q + i);
            v[i] = __fmul_rn((float)(int)(signed char)(a & 255u), d);
            v[i + 1] = __fmul_rn((float)(int)(signed char)(a >> 8), d);
        }
    }
    return make_uint4(umma_pack_h2(v[0], v[1]), umma_pack_h2(v[2], v[3]), umma_pack_h2(v[4], v[5]), umma_pack_h2(v[6], v[7]));
}

// Q4_K, 16-byte aligned rows: the 64 elements k0 .. k0 + 63 of a row are the low and high nibbles of 32 consecutive qs
// bytes (one scale / min pair each): header and scales decoded once, two 16-byte loads, written as the row's 8 chunks.
template <bool SM>
__device__ __forceinline__ void umma_deq64_q4k(const uint8_t* row, int k0, uint8_t* sA, int r) {
    const uint8_t* blk = row + (long long)(k0 >> 8) * 144;
    const int gp = (k0 & 255) >> 6;
    const uint4 hdr = umma_ldw<SM>(reinterpret_cast<const uint4*>(blk));   // d | dmin, scales[12]
    const float d = half_bits_to_float(hdr.x), dmin = half_bits_to_float(hdr.x >> 16);
    const uint8_t sc[12] = {(uint8_t)hdr.y, (uint8_t)(hdr.y >> 8), (uint8_t)(hdr.y >> 16), (uint8_t)(hdr.y >> 24),
                            (uint8_t)hdr.z, (uint8_t)(hdr.z >> 8), (uint8_t)(hdr.z >> 16), (uint8_t)(hdr.z >> 24),
                            (uint8_t)hdr.w, (uint8_t)(hdr.w >> 8), (uint8_t)(hdr.w >> 16), (uint8_t)(hdr.w >> 24)};
    int s1, m1, s2, m2;
    scale_min_k4(sc, 2 * gp, s1, m1);
    scale_min_k4(sc, 2 * gp + 1, s2, m2);
    const float d1 = __fmul_rn(d, (float)s1), mm1 = __fmul_rn(dmin, (float)m1), d2 = __fmul_rn(d, (float)s2), mm2 = __fmul_rn(dmin, (float)m2);
    const uint4* q4 = reinterpret_cast<const uint4*>(blk + 16 + 32 * gp);
    const uint4 qa = umma_ldw<SM>(q4), qb = umma_ldw<SM>(q4 + 1);
    const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int c = 0; c < 4; c++) {   // chunk c: elements 8c..8c+7 (low nibbles), chunk 4 + c: elements 32 + 8c.. (high nibbles)
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t w = qw[2 * c + h];
            lo[2 * h] = umma_pack_h2(__fsub_rn(__fmul_rn(d1, (float)(w & 15)), mm1), __fsub_rn(__fmul_rn(d1, (float)((w >> 8) & 15)), mm1));
            lo[2 * h + 1] = umma_pack_h2(__fsub_rn(__fmul_rn(d1, (float)((w >> 16) & 15)), mm1), __fsub_rn(__fmul_rn(d1, (float)((w >> 24) & 15)), mm1));
            hi[2 * h] = umma_pack_h2(__fsub_rn(__fmul_rn(d2, (float)((w >> 4) & 15)), mm2), __fsub_rn(__fmul_rn(d2, (float)((w >> 12) & 15)), mm2));
            hi[2 * h + 1] = umma_pack_h2(__fsub_rn(__fmul_rn(d2, (float)((w >> 20) & 15)), mm2), __fsub_rn(__fmul_rn(d2, (float)((w >> 28) & 15)), mm2));
        }
        *reinterpret_cast<uint4*>(sA + umma_sw128(r, c)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(sA + umma_sw128(r, 4 + c)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
}

// f32 -> fp16 of a GEMM input (per-op entry point; the prefill kernels write fp16 themselves)
__global__ void umma_to_half_kernel(const float* __restrict__ x, __half* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = f2h_sat(x[i]);
}

struct UmmaParams {
    const uint8_t* w;
    long long row_bytes;
    int type, n_rows, K;
    const __half* x;     // [T][ldx] fp16 (rounded once by the kernel that produced the vector, not by every CTA)
    int ldx, T;
    float* y;            // [T][ldy]: Y[t][j]
    int ldy;
    const float* bias;   // optional [n_rows]
    int accumulate;      // Y += ... (residual connection) instead of Y = ...
    // split-K (small T: too few (rows / 128) x (T / TN) tiles to fill the GPU): blockIdx.z handles K range
    // [z * k_split, (z + 1) * k_split) and stores its partial tile to part[z][t][j] (ld = n_rows); umma_reduce_kernel adds
    // the partials in z order (deterministic) and applies bias / accumulate
    int k_split;         // 0 = no split
    float* part;
    int* err;            // watchdog word (set to 5 if the MMA completion never arrives)
    // gemm_umma2.cuh only: one zeroed counter per (row tile, token tile).  With it the split-K partials are summed by the LAST
    // CTA to finish a tile (in z order, so the result does not depend on which one that is) and umma_reduce_kernel is not launched
    int* tile_cnt;
    // TMA-fed weights: tensor map of the matrix ([n_rows][row_bytes / 4] UINT32, box = 128 rows x raw_pitch bytes = one
    // 256-element block per row); nullptr = the threads read their rows straight from global memory
    const void* tmap;
    int raw_pitch;       // bytes between the rows of a raw tile
    int raw_bytes;       // bytes of 256 elements of one row (144 / 176 / 210 / 272)
    int pdl;             // host: launch with programmatic stream serialization (the persistent kernel and its reduce kernel wait
                         // with griddepcontrol.wait after their prologue; batched decode, engine.cu)
};

template <int TN>
__global__ void __launch_bounds__(128) dequant_gemm_umma_kernel(const UmmaParams p) {
    extern __shared__ __align__(1024) uint8_t umma_smem[];
    __shared__ __align__(8) unsigned long long s_bar[4];   // [0,1]: MMA commits per fp16 stage, [2,3]: TMA raw tiles
    __shared__ uint32_t s_tmem;
    constexpr int kStageBytes = (kUmmaM + TN) * 128;
    uint8_t* base = umma_smem + ((1024u - (umma_smem_u32(umma_smem) & 1023u)) & 1023u);   // the swizzle pattern is on absolute address bits
    uint8_t* sA = base;                       // stage s: [128 weight rows][128 B] at sA + s * kStageBytes
    uint8_t* sB = base + kUmmaM * 128;        //          [TN tokens][128 B]      at sB + s * kStageBytes
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row0 = blockIdx.x * kUmmaM, tok0 = blockIdx.y * TN;
    const uint32_t bar = umma_smem_u32(&s_bar[0]);
    constexpr int kCols = TN < 32 ? 32 : TN;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 16u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 24u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma_smem_u32(&s_tmem)), "n"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(kUmmaM, TN);
    const uint8_t* wrow = p.w + (long long)min(row0 + tid, p.n_rows - 1) * p.row_bytes;
    // Two stages of (A, B) tiles: the MMAs of step k run while the threads dequantise step k + 1; a stage is rewritten
    // only after the commit of the step that last read it (two steps back) has arrived on the stage's mbarrier.
    uint32_t ph0 = 0, ph1 = 0;
    bool alive = true;
    const bool q4_fast = p.type == T_Q4_K && !(((uintptr_t)p.w | (uintptr_t)p.row_bytes) & 15);
    int it = 0;
    const int kb = p.k_split ? (int)blockIdx.z * p.k_split : 0, ke = p.k_split ? min(p.K, kb + p.k_split) : p.K;
    // TMA-fed weights: one box of 128 rows x one 256-element block per row lands in a raw stage while the previous block is
    // dequantised from the other one (the per-thread global reads of a row are 2304 bytes apart: uncoalesced and latency
    // bound, which is what held small-T passes at ~250 GB/s)
    const bool tma = p.tmap != nullptr;
    uint8_t* sRaw = sB + 2 * kStageBytes - kUmmaM * 128;   // after the two fp16 stages
    uint32_t rph0 = 0, rph1 = 0;
    auto raw_issue = [&](int blk) {   // tid 0 only
        const int st = blk & 1;
        const uint32_t rb = bar + 16u + 8u * st;
        const int start = blk * p.raw_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rb), "r"(p.raw_pitch * kUmmaM) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         umma_smem_u32(sRaw + st * kUmmaRawStage)),
                     "l"(p.tmap), "r"((start & ~15) >> 2), "r"(row0), "r"(rb)
                     : "memory");
    };
    if (tma && tid == 0) {
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(p.tmap) : "memory");
        raw_issue(kb >> 8);
    }
    for (int k0 = kb; k0 < ke; k0 += kUmmaK, it++) {
        const int sidx = it & 1;
        uint8_t* tA = sA + sidx * kStageBytes;
        uint8_t* tB = sB + sidx * kStageBytes;
        if (it >= 2 && alive) {
            if (!umma_mbar_wait(bar + 8u * sidx, sidx ? ph1 : ph0)) {
                alive = false;
                if (p.err) atomicExch(p.err, 5);
            }
            if (sidx) ph1 ^= 1u; else ph0 ^= 1u;
        }
        if (tma) {
            const int blk = k0 >> 8, st = blk & 1;
            if ((k0 & 255) == 0) {   // first step of a block: its raw tile must have landed
                if (alive && !umma_mbar_wait(bar + 16u + 8u * st, st ? rph1 : rph0)) {
                    alive = false;
                    if (p.err) atomicExch(p.err, 6);
                }
                if (st) rph1 ^= 1u; else rph0 ^= 1u;
            }
            const uint8_t* rrow = sRaw + st * kUmmaRawStage + tid * p.raw_pitch + ((blk * p.raw_bytes) & 15);
            if (p.type == T_Q4_K) {
                umma_deq64_q4k<true>(rrow, k0 & 255, tA, tid);
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(tA + umma_sw128(tid, c)) = umma_deq8<true>(p.type, rrow, k0 & 255, c);
            }
        } else if (q4_fast) {
            umma_deq64_q4k<false>(wrow, k0, tA, tid);
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(tA + umma_sw128(tid, c)) = umma_deq8<false>(p.type, wrow, k0, c);
        }
        {
            constexpr int NB = TN * 8 / 128;   // 16-byte chunks of the activation tile per thread
            uint4 v[NB];
#pragma unroll
            for (int u = 0; u < NB; u++) {     // all loads first (independent), then the stores
                const int i = tid + u * 128, r = i >> 3, c = i & 7, tk = tok0 + r;
                v[u] = tk < p.T ? __ldg(reinterpret_cast<const uint4*>(p.x + (long long)tk * p.ldx + k0 + 8 * c)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int i = tid + u * 128, r = i >> 3, c = i & 7;
                *reinterpret_cast<uint4*>(tB + umma_sw128(r, c)) = v[u];
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tma && tid == 0 && (k0 & 255) == 0 && k0 + 256 < ke) raw_issue((k0 >> 8) + 1);   // everybody is past the previous block: its stage is free
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = umma_desc(umma_smem_u32(tA)), db = umma_desc(umma_smem_u32(tB));
#pragma unroll
            for (int kk = 0; kk < kUmmaK / 16; kk++)   // 16 fp16 = 32 bytes along the swizzled row: start address + 2
                umma_f16(tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc, (k0 > kb || kk > 0) ? 1u : 0u);
            umma_commit(bar + 8u * sidx);
        }
    }
    // the accumulator is complete when the last commit has arrived (commits arrive in issue order)
    if (it > 0 && alive) {
        const int sidx = (it - 1) & 1;
        if (!umma_mbar_wait(bar + 8u * sidx, sidx ? ph1 : ph0) && p.err) atomicExch(p.err, 5);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 = weight rows, 32 token columns per load ----
    const int j = row0 + tid;
    const float bj = (p.bias && j < p.n_rows) ? p.bias[j] : 0.0f;
#pragma unroll 1
    for (int n0 = 0; n0 < TN; n0 += 32) {
        if (tok0 + n0 >= p.T) break;   // CTA-uniform
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int n = 0; n < 32; n++) {
            const int tk = tok0 + n0 + n;
            if (tk < p.T && j < p.n_rows) {
                if (p.k_split) {
                    p.part[((long long)blockIdx.z * p.T + tk) * p.n_rows + j] = __uint_as_float(v[n]);
                } else {
                    float* yp = p.y + (long long)tk * p.ldy + j;
                    float val = __uint_as_float(v[n]) + bj;
                    if (p.accumulate) val += *yp;
                    *yp = val;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols) : "memory");
}

// Y[t][j] (+)= sum_z part[z][t][j] + bias[j]   (split-K epilogue, z in fixed order)
__global__ void umma_reduce_kernel(const UmmaParams p, int splits) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n = (long long)p.T * p.n_rows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / p.n_rows), j = (int)(i - (long long)t * p.n_rows);
        float acc = 0.0f;
        for (int z = 0; z < splits; z++) acc += p.part[(long long)z * n + i];
        if (p.bias) acc += p.bias[j];
        float* yp = p.y + (long long)t * p.ldy + j;
        *yp = p.accumulate ? *yp + acc : acc;
    }
}

inline bool umma_type_ok(int type) { return type == T_Q4_K || type == T_Q5_K || type == T_Q6_K || type == T_Q8_0; }
// Is the launch eligible?  K a multiple of 256 (K-quants) / 64, rows at least 2-byte aligned (4 for Q4_K / Q5_K headers),
// fp16 activation rows 16-byte aligned.
inline bool umma_eligible(const UmmaParams& p) {
    if (!umma_type_ok(p.type) || p.K <= 0 || p.K % 64 || p.K % type_block_elems(p.type) || p.T <= 0 || p.n_rows <= 0) return false;
    const bool k45 = p.type == T_Q4_K || p.type == T_Q5_K;
    if (((uintptr_t)p.w | (uintptr_t)p.row_bytes) & (k45 ? 7 : 1)) return false;
    if (((uintptr_t)p.x & 15) || (p.ldx & 7)) return false;
    return true;
}
template <int TN>
inline cudaError_t umma_launch_tn(const UmmaParams& p, cudaStream_t st) {
    const size_t smem = (size_t)2 * (kUmmaM + TN) * 128 + (p.tmap ? 2 * kUmmaRawStage : 0) + 1024;   // two fp16 stages (+ two raw stages)
    static bool once[64] = {};   // per instantiation AND device (function attributes are per device): the larger of the two layouts
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !once[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dequant_gemm_umma_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)((size_t)2 * (kUmmaM + TN) * 128 + 2 * kUmmaRawStage + 1024));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) once[dev] = true;
    }
    const int splits = p.k_split ? (p.K + p.k_split - 1) / p.k_split : 1;
    dim3 grid((p.n_rows + kUmmaM - 1) / kUmmaM, (p.T + TN - 1) / TN, splits);
    dequant_gemm_umma_kernel<TN><<<grid, 128, smem, st>>>(p);
    if (splits > 1) {
        const long long n = (long long)p.T * p.n_rows;
        umma_reduce_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>(p, splits);
    }
    return cudaGetLastError();
}
// Split-K plan for a launch with few tiles: K ranges of at least 512 elements (multiples of 256), at most 8, enough to put
// ~2 CTAs on every SM.  `scratch_floats` = capacity of `part`; returns with p.k_split / p.part set (0 = no split).
inline void umma_plan_split(UmmaParams& p, float* scratch, size_t scratch_floats, int n_sm) {
    p.k_split = 0;
    p.part = nullptr;
    if (!scratch || p.T > 64) return;
    const int tn = p.T <= 32 ? 32 : 64;
    const int tiles = ((p.n_rows + kUmmaM - 1) / kUmmaM) * ((p.T + tn - 1) / tn);
    int splits = std::min(8, std::min(p.K / 512, (2 * n_sm + tiles - 1) / tiles));
    while (splits > 1 && (size_t)splits * p.T * p.n_rows > scratch_floats) splits--;
    if (splits <= 1) return;
    int ks = ((p.K + splits - 1) / splits + 255) & ~255;
    if (ks >= p.K) return;
    p.k_split = ks;
    p.part = scratch;
}
// Split-K plan for the PERSISTENT kernel (gemm_umma2.cuh: one CTA per SM walks the items): what counts is the longest chain of
// K steps one CTA executes, rounds(items / n_sm) x steps per item, plus the reduce kernel and the partial-tile traffic a split
// costs (taken as 8 steps).  Llama-3-8B at 32 rows on 148 SMs: q / o (32 tiles x 64 steps) -> 4 ranges (one round of 16 steps),
// k / v (8 tiles) -> 8, down (32 x 224) -> 4 (one round of 56), but gate / up (112 tiles x 64 steps) -> NO split: 3 ranges (what the rule above picks) make 336 items =
// 3 rounds x 24 steps, longer than the 64 steps of the unsplit tiles, and add a reduce launch each.
inline void umma2_plan_split(UmmaParams& p, float* scratch, size_t scratch_floats, int n_sm) {
    p.k_split = 0;
    p.part = nullptr;
    if (!scratch || p.T > 64) return;
    const int tn = p.T <= 32 ? 32 : 64;
    const long long tiles = (long long)((p.n_rows + kUmmaM - 1) / kUmmaM) * ((p.T + tn - 1) / tn);
    const int blocks = p.K / 256;
    long long best_cost = ((tiles + n_sm - 1) / n_sm) * (long long)blocks * 4;
    int best = 1;
    for (int sN = 2; sN <= 8 && sN <= blocks / 2; sN++) {
        if ((size_t)sN * p.T * p.n_rows > scratch_floats) break;
        const int kb = (blocks + sN - 1) / sN;               // blocks per range
        const int n_z = (blocks + kb - 1) / kb;              // ranges actually used
        const long long cost = ((tiles * n_z + n_sm - 1) / n_sm) * (long long)kb * 4 + 8;
        if (cost < best_cost) { best_cost = cost; best = sN; }
    }
    if (best <= 1) return;
    p.k_split = ((blocks + best - 1) / best) * 256;
    p.part = scratch;
}
inline cudaError_t umma_launch(const UmmaParams& p, cudaStream_t st) {
    if (p.T <= 32) return umma_launch_tn<32>(p, st);
    if (p.T <= 64) return umma_launch_tn<64>(p, st);
    if (p.T <= 128) return umma_launch_tn<128>(p, st);
    return umma_launch_tn<256>(p, st);
}

}  // namespace b200
