// gemv_mma.cuh — the production dequant-GEMV for batch-1 decode (Q4_K / Q5_K / Q6_K / Q8_0).
//
// y[j] = sum_k deq(W)[j,k] * x[k] on the untouched GGUF super-block layout (row j of W is
// K/bs contiguous blocks), replacing the reference's vec_mat_q* kernels
// (src/backend/cuda/kernels.rs:443-735) and the CPU hot loop fused_vecmat_dispatch ->
// simd::dot_q* (src/backend/cpu/ops.rs:1123-1191, src/backend/cpu/simd.rs:931-1146).
// Fused around it: RMSNorm of x (simd.rs:847-899), +bias (layers.rs:68-74), +residual
// (layers.rs:1202-1241), silu(gate)*up (simd.rs:598-649), out += w_e*y (moe.rs:363-368).
//
// Second generation (profiles/r01_mega_ncu_full.md: the fp16 HMMA version spent 4.1 thread-instructions per
// weight, half of them outside the dot products, and the token was issue-bound):
//  * the dot products are INTEGER tensor-pipe MMAs (mma.sync.m16n8k32 u8 x s8 -> s32, SASS IMMA.16832):
//    a 32-bit word of nibbles becomes four A-operand quants with ONE LOP3 (w & 0x0F0F0F0F; the high nibbles are
//    w & 0xF0F0F0F0 = 16 q, the 1/16 goes into the scale), i.e. 0.25 instructions per weight for the unpack
//    instead of 0.5 + the fp16 bias handling;
//  * x is staged once per launch as three int8 planes: every 32 elements are scaled by their own power of two so
//    that the largest is in [2^20, 2^21), rounded to an integer and cut into signed bytes hi/mid/lo
//    (x_int = 65536 hi + 256 mid + lo, 21-22 significant bits relative to the group's largest element, the sums
//    are exact in s32).  The 8 columns of an MMA carry (hi, mid) pairs of FOUR sub-blocks, a second accumulator
//    the lo bytes; the B operand is zero except in the columns of the sub-block a lane's k-slots belong to, so
//    each sub-block's sums land in the lane that decodes its scale and nothing is shuffled until a tile is done;
//  * block scales / mins are applied in f32 to per-sub-block sums (the reference's separated form
//    d*sc*sum(q x) - dmin*m*sum(x), simd.rs:1006-1013); 6-bit scales are unpacked for two rows at a time and
//    converted with PRMT into the mantissa of 2^23 (no I2F);
//  * a unit is 32 rows x 256 elements: per-unit overhead (producer, B-operand loads, loop) is paid once per
//    8192 weights, and a finished tile is 32 consecutive outputs in one warp (one x-scale group of the next GEMV);
//  * weights go HBM -> shared memory with 16-byte cp.async (SASS LDGSTS) into PER-WARP rings of 2-4 stages; a warp
//    produces and consumes its own ring (no cross-warp hand-off) and the first stages are issued BEFORE the
//    dependency wait (weights never depend on the previous kernel / phase of the token);
//  * stream-K: all (tile, chunk) units of a launch are dealt evenly, in contiguous runs, to the warps of a
//    persistent grid (one CTA per SM, 8 warps).  Tiles cut across warps are merged through shared memory inside a
//    CTA and through a small global scratch + ticket across CTAs, always in a fixed order (deterministic).
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

constexpr int kMmaMaxWarps = 8;    // warps per CTA (the per-token megakernel uses the same block size)
constexpr int kMmaMaxStages = 4;
constexpr int kMmaRows = 32;       // rows per tile / unit (two m16 MMA row blocks)

enum : int { ME_STORE = 0, ME_RESIDUAL = 1, ME_SWIGLU = 2, ME_SCALED_ACC = 3 };

constexpr int kMmaMaxPeers = 8;
constexpr int kMmaChunk = 256;   // elements of K per unit (one K-quant super-block, eight Q8_0 blocks)

struct MSeg {
    const uint8_t* w;
    float* out;            // f32 output [n_rows]
    const float* bias;     // optional
    long long row_bytes;
    long long expert_stride;
    int type;
    int n_rows;
    int n_tiles;           // ceil(n_rows / 32)
    int unit0;             // first unit of this segment in the launch
    int row_stride;        // pitch of the row slots of a ring stage
    int cb;                // blocks per unit
    int bb;                // block bytes
    int chunk_bytes;       // cb * bb
    int nb_row;            // blocks per row (K / block elems)
    // streamed megakernel (stream.cuh): TMA tensor map of this matrix ([n_rows][row_bytes / elem] tiled in boxes of
    // 32 rows x s_pitch bytes), nullptr when the matrix is not eligible
    const void* tmap;
    int s_pitch;           // bytes between the rows of a ring entry (= inner box bytes)
    int s_elem;            // bytes per tensor-map element (inner coordinates are in these units)
};

struct MParams {
    MSeg seg[3];
    int n_seg;
    int K;
    int chunks;            // ceil(K / 256)
    int units_per_tile;    // chunks (2*chunks for ME_SWIGLU: gate chunks then up chunks)
    int total_units;
    // dealing: CTA b < n_ctas owns `cbase` (+1 for b < crem) consecutive units (tile_mode 0) or whole 32-row tiles
    // (tile_mode 1: no tile straddles two CTAs, so nothing is merged through global memory); its warps split the
    // CTA's units evenly
    int n_ctas, tile_mode, cbase, crem;
    int stages;
    int pf_units;          // megakernel: units per warp (beyond the ring) pulled towards L2 before the barrier wait
    int stage_bytes;       // 32 * max row_stride
    // streamed megakernel: a ring entry is 32 rows x s_C chunks; a tile (x matrix part) is s_ept entries; whole tiles
    // are dealt to CTAs: CTA b (rotated by s_rot) owns s_cbase (+1 for the first s_crem) tiles of the s_tiles logical ones
    int s_C, s_ept, s_parts, s_tiles, s_ncta, s_cbase, s_crem, s_rot;
    // second streamed megakernel (stream2.cuh): s_E = s_tiles * s_parts * s_ept ring entries, dealt to CTAs as equal contiguous
    // ranges; cand: (vocab head, greedy mode) collect argmax candidates in the epilogue instead of storing the logits
    int s_E, cand;
    int s_J, s_R, s_jsh;   // stream2.cuh: jobs per entry along K (1, or 2 = one per chunk) and along the rows (1, or 2 = one per 16-row block); log2(J R)
    int x_bytes;           // bytes of the staged input (x_staged_bytes(K)): what the loader warp's bulk copy moves
    const float* x;        // [K] f32
    // Producer-staged input (per-token megakernel, single GPU): the phase that produced x also wrote its int8 planes,
    // group scales / sums and per-group sums of squares (stage_out32); this launch copies them instead of converting
    // x in every CTA.  The RMSNorm weight is already applied; norm_w != nullptr only says "scale by 1/rms".
    const uint8_t* x_staged;
    // Producer side: also write the staged form of the output vector (stage_K rows in total, times stage_w[j])
    uint8_t* stage_out;
    const float* stage_w;
    int stage_K;
    const float* norm_w;   // optional fused RMSNorm weight [K]
    float eps;
    int epi;
    const float* residual;
    // MoE (expert-resident): the slot-th selected expert
    const int* expert_sel;
    const float* expert_wt;
    int expert_slot;
    int expert_base, expert_count;   // expert parallel: see GemvParams (gemv.cuh)
    // tensor parallel (megakernel only).  Input side: x = sum_r xsum[r * sum_stride + e] (+ x_res[e]) -- the
    // all-reduce of a row-parallel GEMV is finished by its consumer, in rank order on every rank; CTA 0 also
    // stores the summed vector to x_full_out (it is the residual of a later phase).  Output side: the finished
    // rows go to peer_out[r][j] for every rank r (peer memory over NVLink) instead of seg.out.
    const float* xsum;
    int n_sum, sum_stride;
    const float* x_res;
    float* x_full_out;
    float* peer_out[kMmaMaxPeers];
    int n_peer;
    // stream2.cuh, tensor parallel: the all-reduce of this row-parallel GEMV is finished inside the phase.  The epilogue writes its partial
    // tiles as 8-byte (value, epoch) packets (peer_out[r] then points to uint2 slots), and the consumer warps, when their jobs are done, poll
    // the packets of every rank in local memory, add them in rank order + x_res, and write x_full_out and its staged form (stage_out /
    // stage_w / stage_K describe THAT vector; xsum = packet slots of rank 0, sum_stride in packets).  No fence, no flag, no grid barrier
    // between the GEMV and the reduction.
    int ll_red;
    // cross-CTA merge scratch
    float* part;             // [grid][2][2][32]
    unsigned int* tickets;   // [total logical tiles], zero between launches
    int* err;                // device error flag (kept for the watchdog ABI)
    unsigned long long* dbg; // optional [grid*warps][8] globaltimer stamps (lab only)
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a lost copy must never hang the GPU box (sets *err and gives up)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 400000000LL) {  // ~0.2 s
            if (err) atomicExch(err, 1);
            return false;
        }
    }
    return true;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// Shared-memory loads are NOT volatile so that ptxas/nvcc may interleave them with the MMAs of neighbouring
// blocks.  Ordering against the cp.async pipeline comes from data dependencies: every address is derived from an
// opaque token produced after the wait (smem_token), and the unit's results are pinned before the stage is refilled.
__device__ __forceinline__ uint32_t smem_token() {
    uint32_t t;
    asm volatile("mov.u32 %0, 0;" : "=r"(t)::"memory");
    return t;
}
__device__ __forceinline__ void pin4(float (&a)[4]) { asm volatile("" : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3])::"memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    unsigned short v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (uint32_t)v;
}
__device__ __forceinline__ int lds_s8(uint32_t a) {
    int v;
    asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
// 8 bytes at a 2-byte-aligned shared address, whatever the residue: 3 aligned words + funnel shifts
__device__ __forceinline__ void lds_piece8_any(uint32_t a, uint32_t& o0, uint32_t& o1) {
    const uint32_t base = a & ~3u, sh = (a & 3u) << 3;
    const uint32_t w0 = lds32(base), w1 = lds32(base + 4), w2 = lds32(base + 8);
    o0 = __funnelshift_r(w0, w1, sh);
    o1 = __funnelshift_r(w1, w2, sh);
}
// 8 bytes at a shared address whose alignment class is known at compile time (AL = 8, 4, or anything even)
template <int AL>
__device__ __forceinline__ void lds_piece8(uint32_t a, uint32_t& o0, uint32_t& o1) {
    if (AL == 8) {
        const uint2 v = lds64(a);
        o0 = v.x;
        o1 = v.y;
    } else if (AL == 4) {
        o0 = lds32(a);
        o1 = lds32(a + 4);
    } else {
        lds_piece8_any(a, o0, o1);
    }
}

// Integer tensor-pipe MMAs (SASS IMMA.16832): D[16x8] += A[16x32] * B[32x8], s32 accumulators.
// Fragments (lane = 4n + t): a0 = A[n][4t..4t+3], a1 = A[n+8][4t..], a2 = A[n][16+4t..], a3 = A[n+8][16+4t..];
// b0 = B[4t..4t+3][n], b1 = B[16+4t..][n]; d0 = D[n][2t], d1 = D[n][2t+1], d2 = D[n+8][2t], d3 = D[n+8][2t+1].
__device__ __forceinline__ void imma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// first MMA of a chain: C = 0 (no accumulator initialisation instructions)
__device__ __forceinline__ void imma_u8s8_z(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
__device__ __forceinline__ void imma_s8s8_z(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
__device__ __forceinline__ void imma_s8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------- x in shared memory
// Three int8 planes p0 (hi), p1 (mid), p2 (lo): x[e] * 2^k(e/32) rounded to an integer = 65536 p0 + 256 p1 + p2, natural
// element order (4 consecutive elements = one 32-bit word = the four k-slots of an MMA fragment register).
//   sx[i]  = {2^-k, sum of the 32 elements of group i (true units)}   (float2 per 32 elements)
//   x16[i] = -32 * sum of elements 16i..16i+15 (true units; Q6_K's offset)
// The plane bases are 0 / 32 / 64 (mod 128) bytes so that hi / mid / lo loads of one instruction use different banks;
// 256 zero bytes serve as the B operand of lanes whose columns an MMA does not address.
struct XLayout {
    uint32_t p0, p1, p2, sx, x16, ssq, zero, total;
};
__host__ __device__ inline XLayout x_layout(int K) {
    const uint32_t KP = ((uint32_t)K + 127u) & ~127u;
    XLayout L;
    L.p0 = 0;
    L.p1 = KP + 32;
    L.p2 = 2 * KP + 192;
    L.sx = 3 * KP + 384;
    L.x16 = L.sx + (uint32_t)K / 4;                       // K/32 float2
    L.ssq = L.x16 + (uint32_t)K / 4;                      // K/16 floats; then K/32 floats: sum of x^2 per group (staged form only)
    L.zero = ((L.ssq + (uint32_t)K / 8 + 127u) & ~127u) + 96;
    L.total = ((L.zero + 256u + 127u) & ~127u);
    return L;
}
__host__ __device__ inline size_t x_smem_bytes(int K) { return x_layout(K).total; }
// bytes of the staged form of a K-vector in GLOBAL memory (what a producer writes and a consumer copies: no zero page)
__host__ __device__ inline size_t x_staged_bytes(int K) { return (x_layout(K).zero + 15u) & ~15u; }

struct XSmem {
    uint32_t p0, p1, p2, sx, x16, zero;  // shared-space byte addresses
};

// One float4 of the GEMV input: plain x, or (tensor parallel) the rank-ordered sum of the partial vectors + residual
struct XSource {
    const float* x;
    const float* xsum;
    int n_sum, sum_stride;
    const float* x_res;
    float* x_full_out;
};
__device__ __forceinline__ float4 x_fetch4(const XSource& xs, int e) {
    if (xs.n_sum == 0) return *reinterpret_cast<const float4*>(xs.x + e);
    float4 a = *reinterpret_cast<const float4*>(xs.xsum + e);
    for (int r = 1; r < xs.n_sum; r++) {
        const float4 b = *reinterpret_cast<const float4*>(xs.xsum + (size_t)r * xs.sum_stride + e);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (xs.x_res) {
        const float4 b = *reinterpret_cast<const float4*>(xs.x_res + e);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (xs.x_full_out && blockIdx.x == 0) *reinterpret_cast<float4*>(xs.x_full_out + e) = a;
    return a;
}

// x staging from f32, all threads of the CTA, K % 32 == 0, ONE pass and one __syncthreads (by the caller):
//   y = x * w (w: optional RMSNorm weight; the scalar 1/rms is applied to the finished dot products instead of
//   to every element -- (x*inv)*w in the reference, simd.rs:891-892, differs by one rounding);
//   the 8 threads that hold a group of 32 agree on its power of two (largest element -> [2^20, 2^21)), round with
//   the 1.5*2^23 trick (the integer appears in the mantissa), and cut it into three signed bytes by adding
//   0x808080 and flipping the three sign bits; red[warp] = partial sum of x^2 (before the weight).
constexpr int kXRegs = 4;
struct XStage {
    float4 v[kXRegs], w[kXRegs];
};
__device__ __forceinline__ void stage_x_load(XStage& st, const XSource& xs, const float* __restrict__ norm_w, int K, int nthr) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < kXRegs; i++) {
        const int e = (tid + i * nthr) * 4;
        st.v[i] = (e < K) ? x_fetch4(xs, e) : make_float4(0.f, 0.f, 0.f, 0.f);
        st.w[i] = (norm_w && e < K) ? *reinterpret_cast<const float4*>(norm_w + e) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
}
// the scale group's exponent: k such that am * 2^k is in [2^20, 2^21)
__device__ __forceinline__ int x_group_k(float am) {
    int k = 0;
    if (am > 0.0f && am < 3.0e38f) k = min(max(147 - (int)((__float_as_uint(am) >> 23) & 0xFFu), -100), 100);
    return k;
}
__device__ __forceinline__ float split_store4(float4 v, float4 w, int e, uint8_t* xb, const XLayout& L, unsigned mask) {
    const float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
    float am = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    am = fmaxf(am, __shfl_xor_sync(mask, am, 1));
    am = fmaxf(am, __shfl_xor_sync(mask, am, 2));
    am = fmaxf(am, __shfl_xor_sync(mask, am, 4));
    const int k = x_group_k(am);
    const float up = __int_as_float((127 + k) << 23), down = __int_as_float((127 - k) << 23);
    const int i0 = __float_as_int(fmaf(v.x, up, 12582912.0f)) - 0x4B400000, i1 = __float_as_int(fmaf(v.y, up, 12582912.0f)) - 0x4B400000;
    const int i2 = __float_as_int(fmaf(v.z, up, 12582912.0f)) - 0x4B400000, i3 = __float_as_int(fmaf(v.w, up, 12582912.0f)) - 0x4B400000;
    const uint32_t z0 = (uint32_t)(i0 + 0x808080) ^ 0x808080u, z1 = (uint32_t)(i1 + 0x808080) ^ 0x808080u;
    const uint32_t z2 = (uint32_t)(i2 + 0x808080) ^ 0x808080u, z3 = (uint32_t)(i3 + 0x808080) ^ 0x808080u;
    // bytes of z: (lo, mid, hi, 0).  4x3 byte transpose -> one word per plane
    const uint32_t t01 = __byte_perm(z0, z1, 0x5140), t23 = __byte_perm(z2, z3, 0x5140);   // (z0.lo, z1.lo, z0.mid, z1.mid)
    const uint32_t u01 = __byte_perm(z0, z1, 0x0062), u23 = __byte_perm(z2, z3, 0x0062);   // (z0.hi, z1.hi, -, -)
    *reinterpret_cast<uint32_t*>(xb + L.p2 + e) = __byte_perm(t01, t23, 0x5410);
    *reinterpret_cast<uint32_t*>(xb + L.p1 + e) = __byte_perm(t01, t23, 0x7632);
    *reinterpret_cast<uint32_t*>(xb + L.p0 + e) = __byte_perm(u01, u23, 0x5410);
    int s16 = (i0 + i1) + (i2 + i3);   // exact
    s16 += __shfl_xor_sync(mask, s16, 1);
    s16 += __shfl_xor_sync(mask, s16, 2);
    const int s32 = s16 + __shfl_xor_sync(mask, s16, 4);
    if ((threadIdx.x & 3) == 0) *reinterpret_cast<float*>(xb + L.x16 + 4 * (e >> 4)) = -32.0f * ((float)s16 * down);
    if ((threadIdx.x & 7) == 0) *reinterpret_cast<float2*>(xb + L.sx + 8 * (e >> 5)) = make_float2(down, (float)s32 * down);
    return ss;
}
__device__ __forceinline__ void stage_x_finish(const XStage& st, const XSource& xsrc, const float* __restrict__ norm_w, int K,
                                               uint8_t* smem, float* red /*[kMmaMaxWarps]*/, int nthr) {
    const int tid = threadIdx.x;
    const XLayout L = x_layout(K);
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < kXRegs; i++) {
        const int e = (tid + i * nthr) * 4;
        if (e < K) ss += split_store4(st.v[i], st.w[i], e, smem, L, __activemask());
    }
    // the rest in batches of four independent loads (K = 14336 with 256 threads: 14 float4 per thread)
    for (int e0 = (tid + kXRegs * nthr) * 4; e0 < K; e0 += 4 * nthr * 4) {
        float4 v[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e0 + i * nthr * 4;
            v[i] = (e < K) ? x_fetch4(xsrc, e) : make_float4(0.f, 0.f, 0.f, 0.f);
            w[i] = (norm_w && e < K) ? *reinterpret_cast<const float4*>(norm_w + e) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e0 + i * nthr * 4;
            if (e < K) ss += split_store4(v[i], w[i], e, smem, L, __activemask());
        }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    if (tid < 64) reinterpret_cast<uint32_t*>(smem + L.zero)[tid] = 0u;
}
// after the caller's __syncthreads: the factor for finished dot products (1/rms when normalising)
__device__ __forceinline__ float stage_x_unscale(const float* red, bool norm, float eps, int K, int nwarp) {
    float inv = 1.0f;
    if (norm) {
        float tot = 0.0f;
        for (int w = 0; w < nwarp; w++) tot += red[w];
        inv = 1.0f / sqrtf(tot / (float)K + eps);
    }
    return inv;
}

// The staged form of 32 consecutive elements j0..j0+31 (j0 % 32 == 0) of a vector, written to GLOBAL memory by the warp
// that produced them (lane L holds element j0 + L): the same bytes stage_x_finish would write to shared memory for
// x * w (w = 1 without a norm weight), plus the group's sum of x^2 (before the weight).  All 32 lanes must call it.
__device__ __forceinline__ void stage_out32(float val, float w, int j, int K, uint8_t* xg) {
    const XLayout L = x_layout(K);
    const bool valid = j < K;
    const int lane = threadIdx.x & 31;
    const float x = valid ? val : 0.0f;
    const float v = x * w;
    const float am = warp_max(fabsf(v));
    const int k = x_group_k(am);
    const float up = __int_as_float((127 + k) << 23), down = __int_as_float((127 - k) << 23);
    const int i = __float_as_int(fmaf(v, up, 12582912.0f)) - 0x4B400000;
    const uint32_t z = (uint32_t)(i + 0x808080) ^ 0x808080u;
    int s16 = i;
    s16 += __shfl_xor_sync(0xffffffffu, s16, 1);
    s16 += __shfl_xor_sync(0xffffffffu, s16, 2);
    s16 += __shfl_xor_sync(0xffffffffu, s16, 4);
    s16 += __shfl_xor_sync(0xffffffffu, s16, 8);
    const int s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 16);
    const float ssq = warp_sum(x * x);
    if (!valid) return;   // K % 32 == 0: a group is valid as a whole
    xg[L.p2 + j] = (uint8_t)(z & 0xFFu);
    xg[L.p1 + j] = (uint8_t)((z >> 8) & 0xFFu);
    xg[L.p0 + j] = (uint8_t)((z >> 16) & 0xFFu);
    if ((lane & 15) == 0) *reinterpret_cast<float*>(xg + L.x16 + 4 * (j >> 4)) = -32.0f * ((float)s16 * down);
    if (lane == 0) {
        *reinterpret_cast<float2*>(xg + L.sx + 8 * (j >> 5)) = make_float2(down, (float)s32 * down);
        *reinterpret_cast<float*>(xg + L.ssq + 4 * (j >> 5)) = ssq;
    }
}

__device__ __forceinline__ void attn_stage_out(float val, int j, int K, uint8_t* xg) { stage_out32(val, 1.0f, j, K, xg); }

// ---------------------------------------------------------------- per-type unit kernels
// sp = shared address of the stage (row slot r at sp + r*RS; the row's bytes start at +doff), e0 = element index of
// the unit's first element, lane = 4n + t.  A unit is 32 rows (two MMA row blocks rt = 0, 1) x 256 elements; acc[2rt],
// acc[2rt+1] collect rows 16rt + n and 16rt + n + 8 (partial over t: summed when the tile is finished).
//
// LaneB: where this lane finds its B-operand bytes.  A lane feeds column n with the k-slots of A-lane t, and only the
// columns of the sub-block those k-slots belong to may be non-zero: addresses are zero_page + m * (d + e0) with
// m in {0, 1} fixed per lane (one IMAD per address and unit, no selects in the MMA loop).
struct LaneB {
    uint32_t d1, d2;   // (hi or mid plane, lo plane) address of this lane's first element at e0 = 0, minus the zero page
    uint32_t m[6];
    float hs;          // Q4_K: 1/16 for the sub-blocks this lane owns as D-lane when their quants were carried x16
    uint32_t sel_yz, sel_w;
};

// Q4_K / Q5_K (blocks.rs:114-141).  A-lane t reads 16 qs bytes per 64-byte half c of the block: bytes 64c+16t.. ->
// 64-element group gp = 2c + (t>>1) (low nibbles = sub-block 2gp, high = 2gp+1), positions 16(t&1) + 4i + b.
// Word pair ip (i = 2ip, 2ip+1) is one MMA per nibble kind: k-slots of lanes t<2 / t>=2 belong to two different
// sub-blocks.  A1[c] carries (hi, mid) of sub-block 4c + j in column pair j = 2(t>>1) + kind, A2[c] its lo bytes in
// column 2j: D-lane t owns sub-blocks t (c = 0) and 4 + t (c = 1), whose scales it decodes.
__device__ __forceinline__ LaneB lane_b_k45(const XSmem& sm, int n, int t, bool q5) {
    LaneB b{};
    const bool act = (n >> 2) == (t >> 1);
    const int kind = (n >> 1) & 1;
    const uint32_t eo = 64u * (uint32_t)(t >> 1) + 16u * (uint32_t)(t & 1) + 32u * (uint32_t)kind;
    b.d1 = ((n & 1) ? sm.p1 : sm.p0) + eo - sm.zero;
    b.d2 = sm.p2 + eo - sm.zero;
    b.m[0] = (act && kind == 0) ? 1u : 0u;
    b.m[1] = (act && kind == 1) ? 1u : 0u;
    b.m[2] = (b.m[0] && !(n & 1)) ? 1u : 0u;
    b.m[3] = (b.m[1] && !(n & 1)) ? 1u : 0u;
    b.hs = (!q5 && (t & 1)) ? 0.0625f : 1.0f;
    b.sel_yz = (uint32_t)t | ((uint32_t)(4 + t) << 4);
    b.sel_w = b.sel_yz | (b.sel_yz << 8);
    return b;
}
// byte k of v -> float(2^23 + byte): the byte lands in the mantissa of 0x4B000000
__device__ __forceinline__ float byte_magic(uint32_t v, uint32_t sel) { return __uint_as_float(__byte_perm(v, 0x4B000000u, sel)); }

// get_scale_min_k4 (dequant.rs:213-225) for sub-blocks t and 4+t of two rows at once (block headers h0, h1 =
// {d|dmin, scales[12]}): d*sc and dmin*m as the reference computes them (one rounding each).
__device__ __forceinline__ void k4_scales2(const uint4& h0, const uint4& h1, const LaneB& lb, float (&dsc)[2][2], float (&dm)[2][2]) {
    const uint32_t Y = __byte_perm(h0.y, h1.y, lb.sel_yz), Z = __byte_perm(h0.z, h1.z, lb.sel_yz);
    const uint32_t YZ = __byte_perm(Y, Z, 0x5410);              // (y r0, y r1, z r0, z r1): bytes t of scales[0..3], [4..7]
    const uint32_t W = __byte_perm(h0.w, h1.w, lb.sel_w);       // (w r0, w r1, w r0, w r1): byte t of scales[8..11]
    const uint32_t R = YZ & 0x3F3F3F3Fu;                         // (sc_t r0, sc_t r1, m_t r0, m_t r1)
    const uint32_t W2 = (W & 0x00000F0Fu) | ((W >> 4) & 0x0F0F0000u);
    const uint32_t R2 = W2 | ((YZ >> 2) & 0x30303030u);          // (sc_{4+t} r0, r1, m_{4+t} r0, r1)
    const float d0 = half_bits_to_float(h0.x), n0 = half_bits_to_float(h0.x >> 16);
    const float d1 = half_bits_to_float(h1.x), n1 = half_bits_to_float(h1.x >> 16);
    const float d0b = d0 * -8388608.0f, n0b = n0 * -8388608.0f, d1b = d1 * -8388608.0f, n1b = n1 * -8388608.0f;
    dsc[0][0] = fmaf(d0, byte_magic(R, 0x7440), d0b);
    dsc[0][1] = fmaf(d1, byte_magic(R, 0x7441), d1b);
    dm[0][0] = fmaf(n0, byte_magic(R, 0x7442), n0b);
    dm[0][1] = fmaf(n1, byte_magic(R, 0x7443), n1b);
    dsc[1][0] = fmaf(d0, byte_magic(R2, 0x7440), d0b);
    dsc[1][1] = fmaf(d1, byte_magic(R2, 0x7441), d1b);
    dm[1][0] = fmaf(n0, byte_magic(R2, 0x7442), n0b);
    dm[1][1] = fmaf(n1, byte_magic(R2, 0x7443), n1b);
}

template <bool Q5>
__device__ __forceinline__ void unit_k45(uint32_t sp, uint32_t RS, uint32_t e0, const XSmem& sm, const LaneB& lb, int g, int t,
                                         float (&acc)[4]) {
    constexpr uint32_t QS = Q5 ? 48u : 16u;
    const uint32_t t1 = lb.d1 + e0, t2 = lb.d2 + e0;
    const uint32_t a1l = sm.zero + lb.m[0] * t1, a1h = sm.zero + lb.m[1] * t1;
    const uint32_t a2l = sm.zero + lb.m[2] * t2, a2h = sm.zero + lb.m[3] * t2;
    uint4 bl[2], bh[2], cl[2], ch[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        bl[c] = lds128(a1l + 128u * c);
        bh[c] = lds128(a1h + 128u * c);
        cl[c] = lds128(a2l + 128u * c);
        ch[c] = lds128(a2h + 128u * c);
    }
    const uint2 kx0 = lds64(sm.sx + 8u * ((e0 >> 5) + (uint32_t)t)), kx1 = lds64(sm.sx + 8u * ((e0 >> 5) + (uint32_t)t + 4u));
    const float kf[2] = {__uint_as_float(kx0.x) * lb.hs, __uint_as_float(kx1.x) * lb.hs};
    const float XS[2] = {__uint_as_float(kx0.y), __uint_as_float(kx1.y)};
#pragma unroll
    for (int rt = 0; rt < 2; rt++) {
        const uint32_t r0 = sp + (uint32_t)(16 * rt + g) * RS, r1 = r0 + 8u * RS;
        const uint4 h0 = lds128(r0), h1 = lds128(r1);
        uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = qa;
        if (Q5) {
            qa = lds128(r0 + 16u + 16u * (uint32_t)(t & 1));
            qb = lds128(r1 + 16u + 16u * (uint32_t)(t & 1));
        }
        int A1[2][4], A2[2][4];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint4 W0 = lds128(r0 + QS + 64u * c + 16u * (uint32_t)t), W1 = lds128(r1 + QS + 64u * c + 16u * (uint32_t)t);
            const uint32_t w0[4] = {W0.x, W0.y, W0.z, W0.w}, w1[4] = {W1.x, W1.y, W1.z, W1.w};
            const uint32_t ha[4] = {qa.x, qa.y, qa.z, qa.w}, hb[4] = {qb.x, qb.y, qb.z, qb.w};
            const uint32_t bls[4] = {bl[c].x, bl[c].y, bl[c].z, bl[c].w}, bhs[4] = {bh[c].x, bh[c].y, bh[c].z, bh[c].w};
            const uint32_t cls[4] = {cl[c].x, cl[c].y, cl[c].z, cl[c].w}, chs[4] = {ch[c].x, ch[c].y, ch[c].z, ch[c].w};
#pragma unroll
            for (int ip = 0; ip < 2; ip++) {
                uint32_t lo[4], hi[4];
                const uint32_t src[4] = {w0[2 * ip], w1[2 * ip], w0[2 * ip + 1], w1[2 * ip + 1]};   // fragment order a0..a3
                if (!Q5) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        lo[r] = src[r] & 0x0F0F0F0Fu;
                        hi[r] = src[r] & 0xF0F0F0F0u;   // 16 q: the 1/16 is in lb.hs
                    }
                } else {   // 5th bit (dequant.rs:262-315): bit 2gp of qh byte -> low sub-block, bit 2gp+1 -> high sub-block
                    const uint32_t sh = 2u * (2u * c + (uint32_t)(t >> 1));
                    const uint32_t hq[4] = {ha[2 * ip] >> sh, hb[2 * ip] >> sh, ha[2 * ip + 1] >> sh, hb[2 * ip + 1] >> sh};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        lo[r] = (src[r] & 0x0F0F0F0Fu) | ((hq[r] << 4) & 0x10101010u);
                        hi[r] = ((src[r] >> 4) & 0x0F0F0F0Fu) | ((hq[r] << 3) & 0x10101010u);
                    }
                }
                if (ip == 0) {
                    imma_u8s8_z(A1[c], lo[0], lo[1], lo[2], lo[3], bls[0], bls[1]);
                    imma_u8s8_z(A2[c], lo[0], lo[1], lo[2], lo[3], cls[0], cls[1]);
                } else {
                    imma_u8s8(A1[c], lo[0], lo[1], lo[2], lo[3], bls[2], bls[3]);
                    imma_u8s8(A2[c], lo[0], lo[1], lo[2], lo[3], cls[2], cls[3]);
                }
                imma_u8s8(A1[c], hi[0], hi[1], hi[2], hi[3], bhs[2 * ip], bhs[2 * ip + 1]);
                imma_u8s8(A2[c], hi[0], hi[1], hi[2], hi[3], chs[2 * ip], chs[2 * ip + 1]);
            }
        }
        float dsc[2][2], dm[2][2];
        k4_scales2(h0, h1, lb, dsc, dm);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const float s0 = fmaf((float)(A1[c][0] * 256 + A1[c][1]), 256.0f, (float)A2[c][0]);
            const float s1 = fmaf((float)(A1[c][2] * 256 + A1[c][3]), 256.0f, (float)A2[c][2]);
            acc[2 * rt] = fmaf(dsc[c][0] * kf[c], s0, fmaf(-dm[c][0], XS[c], acc[2 * rt]));
            acc[2 * rt + 1] = fmaf(dsc[c][1] * kf[c], s1, fmaf(-dm[c][1], XS[c], acc[2 * rt + 1]));
        }
    }
}

// Q6_K (blocks.rs:143-155, dequant.rs:321-356): ql[128] qh[64] scales[16] d; 16 scale groups of 16 elements.
// Per 128-element half hf, A-lane t reads 8-byte pieces ql[64hf + 8t..], ql[64hf + 32 + 8t..], qh[32hf + 8t..]:
// positions l = 8t..8t+7 of the four quarters q (element 128hf + 32q + l, scale group 8hf + 2q + (l>>4)).  One MMA
// per quarter; its k-slots belong to two scale groups (t<2 / t>=2).  A1[hf][q>>1] carries (hi, mid) of group
// 8hf + 4(q>>1) + j in column pair j = 2(q&1) + (t>>1); A2[hf] the lo bytes of group 8hf + s in column
// 2(s&3) + (s>>2): D-lane t owns groups 8hf + t and 8hf + 4 + t.
__device__ __forceinline__ LaneB lane_b_q6k(const XSmem& sm, int n, int t) {
    LaneB b{};
    const bool act = ((n >> 1) & 1) == (t >> 1);
    b.d1 = ((n & 1) ? sm.p1 : sm.p0) + 8u * (uint32_t)t - sm.zero;
    b.d2 = sm.p2 + 8u * (uint32_t)t - sm.zero;
    b.m[0] = (act && (n >> 2) == 0) ? 1u : 0u;   // even quarters
    b.m[1] = (act && (n >> 2) == 1) ? 1u : 0u;   // odd quarters
    const int q2 = (n >> 2) + 2 * (n & 1);
#pragma unroll
    for (int q = 0; q < 4; q++) b.m[2 + q] = (act && q == q2) ? 1u : 0u;
    b.hs = 1.0f;
    return b;
}
template <int AL>
__device__ __forceinline__ void unit_q6k(uint32_t sp, uint32_t RS, uint32_t e0, uint32_t doff, const XSmem& sm, const LaneB& lb,
                                         int g, int t, float (&acc)[4]) {
    const uint32_t t1 = lb.d1 + e0, t2 = lb.d2 + e0;
    const uint32_t a1[2] = {sm.zero + lb.m[0] * t1, sm.zero + lb.m[1] * t1};
    uint2 B1[2][4], B2[2][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t a2 = sm.zero + lb.m[2 + q] * t2;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            B1[hf][q] = lds64(a1[q & 1] + 128u * hf + 32u * q);
            B2[hf][q] = lds64(a2 + 128u * hf + 32u * q);
        }
    }
#pragma unroll 1
    for (int rt = 0; rt < 2; rt++) {
        const uint32_t blk0 = sp + (uint32_t)(16 * rt + g) * RS + doff, blk1 = blk0 + 8u * RS;
        int A1[2][2][4], A2[2][4];
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            uint32_t QA[4], QB[4], QH[4];   // fragment order: (row n word 0, row n+8 word 0, row n word 1, row n+8 word 1)
            lds_piece8<AL>(blk0 + 64u * hf + 8u * (uint32_t)t, QA[0], QA[2]);
            lds_piece8<AL>(blk1 + 64u * hf + 8u * (uint32_t)t, QA[1], QA[3]);
            lds_piece8<AL>(blk0 + 64u * hf + 32u + 8u * (uint32_t)t, QB[0], QB[2]);
            lds_piece8<AL>(blk1 + 64u * hf + 32u + 8u * (uint32_t)t, QB[1], QB[3]);
            lds_piece8<AL>(blk0 + 128u + 32u * hf + 8u * (uint32_t)t, QH[0], QH[2]);
            lds_piece8<AL>(blk1 + 128u + 32u * hf + 8u * (uint32_t)t, QH[1], QH[3]);
            uint32_t a[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                a[0][r] = (QA[r] & 0x0F0F0F0Fu) | ((QH[r] << 4) & 0x30303030u);
                a[1][r] = (QB[r] & 0x0F0F0F0Fu) | ((QH[r] << 2) & 0x30303030u);
                a[2][r] = ((QA[r] >> 4) & 0x0F0F0F0Fu) | (QH[r] & 0x30303030u);
                a[3][r] = ((QB[r] >> 4) & 0x0F0F0F0Fu) | ((QH[r] >> 2) & 0x30303030u);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((q & 1) == 0) imma_u8s8_z(A1[hf][q >> 1], a[q][0], a[q][1], a[q][2], a[q][3], B1[hf][q].x, B1[hf][q].y);
                else imma_u8s8(A1[hf][q >> 1], a[q][0], a[q][1], a[q][2], a[q][3], B1[hf][q].x, B1[hf][q].y);
                if (q == 0) imma_u8s8_z(A2[hf], a[q][0], a[q][1], a[q][2], a[q][3], B2[hf][q].x, B2[hf][q].y);
                else imma_u8s8(A2[hf], a[q][0], a[q][1], a[q][2], a[q][3], B2[hf][q].x, B2[hf][q].y);
            }
        }
        const float d0 = half_bits_to_float(lds16(blk0 + 208u)), d1 = half_bits_to_float(lds16(blk1 + 208u));
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int hf = 0; hf < 2; hf++)
#pragma unroll
            for (int qq = 0; qq < 2; qq++) {
                const uint32_t sg = 8u * hf + 4u * qq + (uint32_t)t;
                const float s0 = fmaf((float)(A1[hf][qq][0] * 256 + A1[hf][qq][1]), 256.0f, (float)A2[hf][qq]);
                const float s1 = fmaf((float)(A1[hf][qq][2] * 256 + A1[hf][qq][3]), 256.0f, (float)A2[hf][2 + qq]);
                const float kf = lds_f32(sm.sx + 8u * ((e0 >> 5) + (sg >> 1)));
                const float xn = lds_f32(sm.x16 + 4u * ((e0 >> 4) + sg));   // -32 * sum(x) of the group
                const float sc0 = (float)lds_s8(blk0 + 192u + sg), sc1 = (float)lds_s8(blk1 + 192u + sg);
                r0 = fmaf(d0 * sc0, fmaf(kf, s0, xn), r0);
                r1 = fmaf(d1 * sc1, fmaf(kf, s1, xn), r1);
            }
        if (rt == 0) { acc[0] += r0; acc[1] += r1; } else { acc[2] += r0; acc[3] += r1; }
    }
}

// Q8_0 (blocks.rs:60-70): 34-byte blocks of 32, signed quants (s8 x s8 MMA, no offset).  A-lane t reads bytes
// 8t..8t+7 of a block: one MMA per block.  A1[b>>2] carries (hi, mid) of block b in column pair b&3, A2 the lo bytes
// of block b in column 2(b&3) + (b>>2): D-lane t owns blocks t and 4 + t.
__device__ __forceinline__ LaneB lane_b_q80(const XSmem& sm, int n, int t) {
    LaneB b{};
    b.d1 = ((n & 1) ? sm.p1 : sm.p0) + 8u * (uint32_t)t - sm.zero;
    b.d2 = sm.p2 + 8u * (uint32_t)t - sm.zero;
    b.m[0] = (uint32_t)(n >> 1);                 // block (mod 4) this lane feeds in A1
    b.m[1] = (uint32_t)((n >> 1) + 4 * (n & 1)); // block this lane feeds in A2
    b.hs = 1.0f;
    return b;
}
__device__ __forceinline__ void unit_q80(uint32_t sp, uint32_t RS, uint32_t e0, uint32_t doff, int nblk, const XSmem& sm,
                                         const LaneB& lb, int g, int t, float (&acc)[4]) {
    const uint32_t t1 = sm.zero + lb.d1 + e0, t2 = sm.zero + lb.d2 + e0;
    uint2 B1[8], B2[8];
#pragma unroll
    for (int b = 0; b < 8; b++) {
        B1[b] = lds64(((uint32_t)(b & 3) == lb.m[0]) ? t1 + 32u * b : sm.zero);
        B2[b] = lds64(((uint32_t)b == lb.m[1]) ? t2 + 32u * b : sm.zero);
    }
#pragma unroll 1
    for (int rt = 0; rt < 2; rt++) {
        const uint32_t row0 = sp + (uint32_t)(16 * rt + g) * RS + doff, row1 = row0 + 8u * RS;
        int A1[2][4], A2[4];
#pragma unroll
        for (int i = 0; i < 4; i++) A1[0][i] = A1[1][i] = A2[i] = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            if (b < nblk) {   // warp-uniform (ragged last chunk)
                uint32_t a0, a1, a2, a3;
                lds_piece8_any(row0 + 34u * b + 2u + 8u * (uint32_t)t, a0, a2);
                lds_piece8_any(row1 + 34u * b + 2u + 8u * (uint32_t)t, a1, a3);
                imma_s8s8(A1[b >> 2], a0, a1, a2, a3, B1[b].x, B1[b].y);
                imma_s8s8(A2, a0, a1, a2, a3, B2[b].x, B2[b].y);
            }
        }
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int b = t + 4 * j;
            if (b < nblk) {
                const float s0 = fmaf((float)(A1[j][0] * 256 + A1[j][1]), 256.0f, (float)A2[j]);
                const float s1 = fmaf((float)(A1[j][2] * 256 + A1[j][3]), 256.0f, (float)A2[2 + j]);
                const float kf = lds_f32(sm.sx + 8u * ((e0 >> 5) + (uint32_t)b));
                r0 = fmaf(half_bits_to_float(lds16(row0 + 34u * b)) * kf, s0, r0);
                r1 = fmaf(half_bits_to_float(lds16(row1 + 34u * b)) * kf, s1, r1);
            }
        }
        if (rt == 0) { acc[0] += r0; acc[1] += r1; } else { acc[2] += r0; acc[3] += r1; }
    }
}

// ---------------------------------------------------------------- the kernel
__device__ __forceinline__ float mma_silu(float x) { return x / (1.0f + expf(-x)); }

__host__ __device__ inline int mma_chunk_blocks(int type, int chunk_elems) { return chunk_elems / type_block_elems(type); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {  // at most N newest groups still in flight
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {  // warp-uniform n in 0..3
    if (n <= 0) cp_async_wait<0>();
    else if (n == 1) cp_async_wait<1>();
    else if (n == 2) cp_async_wait<2>();
    else cp_async_wait<3>();
}
// ticket with release (my partial sums are visible) + acquire (I see the others') semantics
__device__ __forceinline__ unsigned int atom_add_acq_rel(unsigned int* p, unsigned int v) {
    unsigned int r;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ float ld_relaxed_gpu(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MMA_STAMP(i)                                                    \
    do {                                                                \
        if (p.dbg && lane == 0) p.dbg[(size_t)gw * 8 + (i)] = gtimer(); \
    } while (0)

// Position of a warp in its run of units: segment s, 32-row tile, matrix (ME_SWIGLU: 0 = gate, 1 = up; else = s),
// chunk (256 elements) within the row.
struct MCursor {
    int s, tile, mat, chunk;
};

// The whole GEMV of one launch (or of one phase of the per-token megakernel, mega.cuh) for this CTA.
//   smem   : dynamic shared memory (x staging + rings), 128-byte aligned
//   s_red  : [2 * kMmaMaxWarps] floats, s_part: [kMmaMaxWarps][2][2][32] floats (static shared memory of the caller)
//   pre()  : runs first (megakernel: __syncthreads + arrive at the grid barrier that ends the previous phase)
//   post() : runs after the prologue, before anything that depends on other CTAs / the previous kernel
//            (megakernel: wait at that barrier; stand-alone kernel: griddepcontrol)
//   early  : issue the first ring stages with cp.async BEFORE post() (weights never depend on a predecessor), so
//            they land while the barrier is being waited for; otherwise they are issued after the x loads
//   warm_l2: (not early) pull the first stages towards L2 before post()
template <class Pre, class Post>
__device__ __forceinline__ void mma_gemv_cta(const MParams& p, uint8_t* smem, float* s_red, float (*s_part)[2][2][32], Pre pre_fn,
                                             Post post_fn, bool early, bool warm_l2) {
    pre_fn();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nw = blockDim.x >> 5;
    const int K = p.K;
    const int STAGES = p.stages;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ring = sbase + (uint32_t)x_smem_bytes(K) + (uint32_t)warp * STAGES * p.stage_bytes;
    const bool swiglu = p.epi == ME_SWIGLU;

    // units are dealt in contiguous runs, CTA after CTA, warp after warp: the pieces of a tile meet inside one CTA
    // (merged through shared memory); only tiles that straddle CTAs (tile_mode 0) go through global memory
    const int upt = p.units_per_tile;
    const int cta = blockIdx.x;
    const long long gw = (long long)blockIdx.x * nw + warp;  // debug stamps only
    int cu0 = 0, cu1 = 0;
    if (cta < p.n_ctas) {
        const int first = cta * p.cbase + min(cta, p.crem), cnt = p.cbase + (cta < p.crem ? 1 : 0);
        cu0 = p.tile_mode ? first * upt : first;
        cu1 = p.tile_mode ? (first + cnt) * upt : first + cnt;
    }
    const int cn = cu1 - cu0;
    const int u0 = cu0 + (warp * cn) / nw, u1 = cu0 + ((warp + 1) * cn) / nw;
    const int n_units = u1 - u0;
    long long eoff = 0;  // MoE expert index (valid after post_fn)
    MMA_STAMP(0);

    auto cur_init = [&](MCursor& q, int u) {
        q.s = (p.n_seg > 2 && u >= p.seg[2].unit0) ? 2 : (p.n_seg > 1 && !swiglu && u >= p.seg[1].unit0) ? 1 : 0;
        const int local = u - p.seg[q.s].unit0;
        q.tile = local / p.units_per_tile;
        q.chunk = local - q.tile * p.units_per_tile;
        q.mat = q.s;
        if (swiglu && q.chunk >= p.chunks) { q.mat = 1; q.chunk -= p.chunks; }
    };
    // the cursor has just moved past the last chunk of its row: next matrix (SwiGLU up rows) / tile / segment
    auto cur_wrap = [&](MCursor& q) {
        q.chunk = 0;
        if (swiglu && q.mat == 0) {
            q.mat = 1;
        } else {
            q.tile++;
            if (swiglu) {
                q.mat = 0;
            } else {
                if (q.tile == p.seg[q.s].n_tiles && q.s + 1 < p.n_seg) { q.s++; q.tile = 0; }
                q.mat = q.s;
            }
        }
    };

    // ---- producer: cp.async the 32 rows of a unit into a ring stage.  Lane (g, t) moves the 16-byte pieces
    // t, t+4, ... of rows g, g+8, g+16, g+24 (8 rows x 64 contiguous bytes per instruction); sources are aligned down
    // to 16 bytes, the residue (doff) is re-derived by the consumer from the same address. ----
    MCursor cp{};
    const uint8_t* psrc[4] = {nullptr, nullptr, nullptr, nullptr};   // this lane's first 16-byte piece of each of its rows
    int p_da = 0, p_short = 0, p_cbytes = 0;   // p_da: residue mod 16 of the unit's first byte in this lane's rows
    uint32_t p_dst = 0, p_rs8 = 0;   // this lane's first destination in stage 0, 8 row slots further
    auto prod_run = [&]() {  // per-run constants of the producer's matrix and tile
        const MSeg& sg = p.seg[cp.mat];
        p_cbytes = sg.chunk_bytes;
        p_short = sg.chunk_bytes - (sg.nb_row - (p.chunks - 1) * sg.cb) * sg.bb;   // bytes the last chunk of a row is short of
        p_dst = ring + (uint32_t)g * sg.row_stride + 16u * t;
        p_rs8 = 8u * sg.row_stride;
        const uint8_t* base = sg.w + eoff * sg.expert_stride + (long long)cp.chunk * sg.chunk_bytes;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint8_t* row = base + (long long)min(cp.tile * kMmaRows + g + 8 * j, sg.n_rows - 1) * sg.row_bytes;
            const uint32_t da = (uint32_t)((uintptr_t)row & 15u);
            psrc[j] = row - da + 16 * t;
            // bytes from this lane's first piece to the end of the row's unit.  The four rows of a lane have the same
            // residue (8 * row_bytes is a multiple of 16); rows clamped at the ragged end of a matrix may differ by a
            // few bytes -- their results are discarded and every tensor has 256 bytes of slack
            if (j == 0) p_da = (int)da;
        }
    };
    auto issue = [&](uint32_t stage_off, bool l2_only) {
        const int ea = p_da + p_cbytes - 16 * t - ((cp.chunk == p.chunks - 1) ? p_short : 0);
        const uint32_t dst = p_dst + stage_off;
        if (l2_only) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (112 * t < ea) prefetch_l2(psrc[j] + 112 * t);
        } else {
#pragma unroll
            for (int i = 0; i < 5; i++) {   // <= 287 bytes per row and unit
                if (64 * i < ea) {
#pragma unroll
                    for (int j = 0; j < 4; j++) cp_async16(dst + (uint32_t)j * p_rs8 + 64u * i, psrc[j] + 64 * i);
                }
            }
        }
        cp.chunk++;
        if (cp.chunk == p.chunks) {
            cur_wrap(cp);
            prod_run();
        } else {   // the sources stay 16-byte aligned: advance by whole pieces, carry the residue
            const int nd = p_da + p_cbytes, adv = nd & ~15;
            p_da = nd & 15;
#pragma unroll
            for (int j = 0; j < 4; j++) psrc[j] += adv;
        }
    };

    MCursor cc{};
    const int pre = min(STAGES - 1, n_units);
    if (n_units > 0 && !p.expert_sel) {
        // dense weights never depend on a predecessor: start (or at least pull towards L2) the first stages before
        // the dependency wait
        cur_init(cp, u0);
        cc = cp;
        prod_run();
        if (early) {
            for (int k = 0; k < STAGES - 1; k++) {
                if (k < pre) issue((uint32_t)k * p.stage_bytes, false);
                cp_async_commit();
            }
        } else if (warm_l2) {
            for (int k = 0; k < pre; k++) issue(0, true);
            cp = cc;
            prod_run();
        }
    } else if (early && !p.expert_sel) {
        for (int k = 0; k < STAGES - 1; k++) cp_async_commit();
    }
    const bool issued_early = early && !p.expert_sel;

    // While the CTA waits for its predecessors (grid barrier of the megakernel), its warps pull the units after the
    // ring stages towards L2, one instruction (lane = row: 32 lines of 128 bytes) per call, nearest first; the caller
    // stops calling as soon as the barrier opens, so a late CTA never delays itself (HBM keeps streaming this phase's
    // weights during the barrier instead of idling).
    MCursor pfq = cp;
    int pf_left = (issued_early && n_units > 0) ? min(p.pf_units, n_units - pre) : 0, pf_off = 0, pf_bytes = -1;
    const uint8_t* pf_a = nullptr;
    auto prefetch_step = [&]() -> bool {
        if (pf_left <= 0) return false;
        if (pf_bytes < 0) {   // next run of the cursor
            const int len = min(p.chunks - pfq.chunk, pf_left);
            const MSeg& sg = p.seg[pfq.mat];
            const uint8_t* a = sg.w + (long long)min(pfq.tile * kMmaRows + lane, sg.n_rows - 1) * sg.row_bytes + (long long)pfq.chunk * sg.chunk_bytes;
            pf_bytes = len * sg.chunk_bytes + (int)((uintptr_t)a & 127u);
            pf_a = a - ((uintptr_t)a & 127u);
            pf_off = 0;
            pf_left -= len;
            if (pfq.chunk + len == p.chunks) cur_wrap(pfq); else pfq.chunk += len;
        }
        prefetch_l2(pf_a + pf_off);
        pf_off += 128;
        if (pf_off >= pf_bytes) pf_bytes = -1;
        return true;
    };

    post_fn(prefetch_step);
    MMA_STAMP(1);

    if (p.expert_sel) {
        eoff = (long long)p.expert_sel[p.expert_slot];
        if (p.expert_count > 0) {   // expert parallel: not this GPU's expert -> nothing to do (uniform over the grid)
            eoff -= p.expert_base;
            if (eoff < 0 || eoff >= p.expert_count) return;
        }
        if (n_units > 0) { cur_init(cp, u0); cc = cp; prod_run(); }
    }
    const XLayout XL = x_layout(K);
    float unscale = 1.0f;
    if (p.x_staged) {
        // the producer of x left its staged form in global memory: a flat copy, no arithmetic
        if (!issued_early) {
            for (int k = 0; k < STAGES - 1; k++) {
                if (k < pre) issue((uint32_t)k * p.stage_bytes, false);
                cp_async_commit();
            }
        }
        const uint32_t n16 = (XL.zero + 15u) >> 4;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) cp_async16(sbase + 16u * i, p.x_staged + 16u * i);
        cp_async_commit();
        MMA_STAMP(2);
        cp_async_wait<0>();
        if (threadIdx.x < 64) reinterpret_cast<uint32_t*>(smem + XL.zero)[threadIdx.x] = 0u;
        __syncthreads();
        if (p.norm_w) {   // sum of x^2 from the per-group partial sums, in the same order in every warp and CTA
            const float* ssq = reinterpret_cast<const float*>(smem + XL.ssq);
            float tot = 0.0f;
            for (int i = lane; i < (K >> 5); i += 32) tot += ssq[i];
            tot = warp_sum(tot);
            unscale = 1.0f / sqrtf(tot / (float)K + p.eps);
        }
    } else {
        // x: the loads are issued BEFORE the weight copies, the split after them
        XStage xst;
        const XSource xsrc{p.x, p.xsum, p.n_sum, p.sum_stride, p.x_res, p.x_full_out};
        stage_x_load(xst, xsrc, p.norm_w, K, (int)blockDim.x);
        if (!issued_early) {
            for (int k = 0; k < STAGES - 1; k++) {
                if (k < pre) issue((uint32_t)k * p.stage_bytes, false);
                cp_async_commit();
            }
        }
        MMA_STAMP(2);
        stage_x_finish(xst, xsrc, p.norm_w, K, smem, s_red, (int)blockDim.x);
        __syncthreads();
        unscale = stage_x_unscale(s_red, p.norm_w != nullptr, p.eps, K, (int)(blockDim.x >> 5));
    }
    MMA_STAMP(3);
    const uint32_t tokx = smem_token();
    XSmem sm;
    sm.p0 = sbase + tokx + XL.p0;
    sm.p1 = sbase + tokx + XL.p1;
    sm.p2 = sbase + tokx + XL.p2;
    sm.sx = sbase + tokx + XL.sx;
    sm.x16 = sbase + tokx + XL.x16;
    sm.zero = sbase + tokx + XL.zero;

    // ---- epilogue of a finished tile: lane L owns row tile*32 + L of segment s (vu: the up row for SwiGLU) ----
    auto epilogue = [&](int s, int tile, float v, float vu) {
        const MSeg& sg = p.seg[s];
        const int j = tile * kMmaRows + lane;
        const bool valid = j < sg.n_rows;
        // one round trip for everything the row needs from global memory
        const float e_bias = (valid && sg.bias) ? sg.bias[j] : 0.0f;
        const float e_res = (valid && p.epi == ME_RESIDUAL) ? p.residual[j] : 0.0f;
        const float e_w = (p.stage_out && p.stage_w && j < p.stage_K) ? p.stage_w[j] : 1.0f;
        v *= unscale;
        float val = v;
        if (swiglu) val = mma_silu(v) * (vu * unscale);
        if (valid) {
            val += e_bias;
            val += e_res;
            if (p.epi == ME_SCALED_ACC) {  // moe.rs:363-368
                const float prev = (p.expert_slot == 0 || p.expert_count > 0) ? 0.0f : sg.out[j];
                val = prev + p.expert_wt[p.expert_slot] * val;
                if (p.residual && p.expert_count == 0) val += p.residual[j];
            }
            if (p.n_peer > 0) {
                for (int r = 0; r < p.n_peer; r++) p.peer_out[r][j] = val;  // partial of a row-parallel GEMV, to every rank
            } else {
                sg.out[j] = val;
            }
        }
        if (p.stage_out) stage_out32(val, e_w, j, p.stage_K, p.stage_out);   // the next GEMV's input, ready to copy
    };

    float ag[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f};
    int piece_s[2] = {-1, -1}, piece_tile[2] = {0, 0};  // tiles of which this warp holds only a piece (first / last of its run)
    uint32_t off_c = 0, off_p = (uint32_t)(STAGES - 1) * p.stage_bytes;  // consumer / producer stage offsets
    const uint32_t ring_bytes = (uint32_t)STAGES * p.stage_bytes;
    int issued = pre, done = 0;
    bool first = true;
    while (done < n_units) {
        // ---- a run: the consecutive units of this warp inside one row-tile of one matrix ----
        const int len = min(p.chunks - cc.chunk, n_units - done);
        const MSeg& wsg = p.seg[cc.mat];
        const int type = wsg.type;
        const uint32_t RS = (uint32_t)wsg.row_stride;
        const uint32_t cbytes = (uint32_t)wsg.chunk_bytes;
        uint32_t e0 = (uint32_t)cc.chunk * kMmaChunk;
        // low address bits of this lane's first row of the unit: the source misalignment (identical for its four rows,
        // 8 * row_bytes being a multiple of 16)
        uint32_t ca = (uint32_t)(uintptr_t)wsg.w + (uint32_t)eoff * (uint32_t)wsg.expert_stride + (uint32_t)cc.chunk * cbytes +
                      (uint32_t)min(cc.tile * kMmaRows + g, wsg.n_rows - 1) * (uint32_t)wsg.row_bytes;
        const LaneB lb = (type == T_Q6_K) ? lane_b_q6k(sm, g, t) : (type == T_Q8_0) ? lane_b_q80(sm, g, t) : lane_b_k45(sm, g, t, type == T_Q5_K);
        // Q6_K: alignment class of the block starts of this run (warp-uniform: rows and tiles differ by multiples of row_bytes)
        const uint32_t rbm = (uint32_t)wsg.row_bytes | (uint32_t)wsg.expert_stride | (uint32_t)(uintptr_t)wsg.w;
        float racc[4] = {0.f, 0.f, 0.f, 0.f};
        // one unit: keep the ring full, wait for the oldest stage, consume it
#define MMA_UNIT(CALL)                                                             \
    for (int i = 0; i < len; i++) {                                                \
        if (issued < n_units) { issue(off_p, false); issued++; }                   \
        off_p = (off_p + p.stage_bytes == ring_bytes) ? 0u : off_p + p.stage_bytes; \
        cp_async_commit();                                                         \
        cp_async_wait_dyn(STAGES - 1);                                             \
        __syncwarp();                                                              \
        if (first) { MMA_STAMP(4); first = false; }                                \
        const uint32_t sp = ring + off_c + smem_token();                           \
        off_c = (off_c + p.stage_bytes == ring_bytes) ? 0u : off_c + p.stage_bytes; \
        float ua[4] = {0.f, 0.f, 0.f, 0.f};                                        \
        CALL;                                                                      \
        pin4(ua); /* the unit's shared-memory reads are complete before the stage can be refilled */ \
        racc[0] += ua[0]; racc[1] += ua[1]; racc[2] += ua[2]; racc[3] += ua[3];    \
        e0 += kMmaChunk;                                                           \
        ca += cbytes;                                                              \
        __syncwarp(); /* every lane is done with this stage */                     \
    }
        switch (type) {
            case T_Q4_K: MMA_UNIT(unit_k45<false>(sp, RS, e0, sm, lb, g, t, ua)) break;
#ifndef B200_LEAN_TEST
            case T_Q5_K: MMA_UNIT(unit_k45<true>(sp, RS, e0, sm, lb, g, t, ua)) break;
#endif
            case T_Q6_K:
                // block b of a row starts at 210 b: 8-byte aligned for b % 4 == 0, 4-byte for b % 4 == 2, else 2-byte (when
                // the rows themselves are 8-byte aligned; the generic variant takes any even residue)
                MMA_UNIT({
                    const uint32_t dof = ca & 15u;
                    const uint32_t cls = (rbm & 7u) ? 1u : (dof & 7u);
                    if (cls == 0u) unit_q6k<8>(sp, RS, e0, dof, sm, lb, g, t, ua);
                    else if (cls == 4u) unit_q6k<4>(sp, RS, e0, dof, sm, lb, g, t, ua);
                    else unit_q6k<2>(sp, RS, e0, dof, sm, lb, g, t, ua);
                })
                break;
#ifndef B200_LEAN_TEST
            default: MMA_UNIT(unit_q80(sp, RS, e0, ca & 15u, min(wsg.cb, wsg.nb_row - (int)(e0 >> 5)), sm, lb, g, t, ua)) break;
#else
            default: break;
#endif
        }
#undef MMA_UNIT
        done += len;
        if (swiglu && cc.mat == 1) {
#pragma unroll
            for (int j = 0; j < 4; j++) au[j] += racc[j];
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) ag[j] += racc[j];
        }

        // ---- tile finished (for this warp)? ----
        const int s = cc.s, tile = cc.tile;
        const bool row_end = (cc.chunk + len == p.chunks);
        const bool tile_done = (row_end && (!swiglu || cc.mat == 1)) || (done == n_units);
        if (row_end) cur_wrap(cc); else cc.chunk += len;
        if (!tile_done) continue;
        if (done == n_units) MMA_STAMP(5);

        // reduce the 4 lanes of a row group, then lane L holds logical row L = 8j + n (register j of lanes 4n..4n+3)
        float vg = 0.f, vu = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], 1);
            ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], 2);
            const float x = __shfl_sync(0xffffffffu, ag[j], 4 * (lane & 7));
            if ((lane >> 3) == j) vg = x;
            ag[j] = 0.f;
        }
        if (swiglu) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                au[j] += __shfl_xor_sync(0xffffffffu, au[j], 1);
                au[j] += __shfl_xor_sync(0xffffffffu, au[j], 2);
                const float x = __shfl_sync(0xffffffffu, au[j], 4 * (lane & 7));
                if ((lane >> 3) == j) vu = x;
                au[j] = 0.f;
            }
        }

        const int tu0 = p.seg[s].unit0 + tile * upt;
        if (u0 <= tu0 && tu0 + upt <= u1) {
            epilogue(s, tile, vg, vu);  // the whole tile is mine
        } else {
            const int slot = (u0 >= tu0) ? 0 : 1;  // tile is my first (slot 0) or starts inside my run (slot 1)
            s_part[warp][slot][0][lane] = vg;
            s_part[warp][slot][1][lane] = vu;
            piece_s[slot] = s;
            piece_tile[slot] = tile;
        }
    }

    // ---- merge the pieces: inside the CTA through shared memory, across CTAs through global memory + ticket ----
    __syncthreads();
    auto warp_of = [&](int u) { return ((u - cu0 + 1) * nw - 1) / cn; };          // local warp that owns unit u of this CTA
    auto cta_first = [&](int c) { return c * p.cbase + min(c, p.crem); };          // first unit of CTA c (tile_mode 0)
    auto cta_of = [&](int u) {                                                     // CTA that owns unit u (tile_mode 0)
        const int big = p.crem * (p.cbase + 1);
        return u < big ? u / (p.cbase + 1) : p.crem + (u - big) / p.cbase;
    };
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
        if (piece_s[slot] < 0) continue;  // warp-uniform
        const int s = piece_s[slot], tile = piece_tile[slot];
        const int tu0 = p.seg[s].unit0 + tile * upt;
        const int lo_u = max(tu0, cu0), hi_u = min(tu0 + upt, cu1) - 1;
        const int lo = warp_of(lo_u), hi = warp_of(hi_u);
        if (warp != lo) continue;  // the first warp of the CTA that holds a piece finishes the tile
        float v = 0.f, vu = 0.f;
        for (int w = lo; w <= hi; w++) {  // fixed order: deterministic
            const int wu0 = cu0 + (w * cn) / nw, wu1 = cu0 + ((w + 1) * cn) / nw;
            if (wu1 == wu0) continue;     // a warp without units holds no piece
            const int ws = (wu0 >= tu0) ? 0 : 1;
            v += s_part[w][ws][0][lane];
            vu += s_part[w][ws][1][lane];
        }
        if (tu0 < cu0 || tu0 + upt > cu1) {  // the tile straddles CTAs (tile_mode 0 only)
            const int tile_id = (s == 0 ? 0 : (s == 1 ? p.seg[0].n_tiles : p.seg[0].n_tiles + p.seg[1].n_tiles)) + tile;
            const int c_first = cta_of(tu0), c_last = cta_of(tu0 + upt - 1);
            float* mine = p.part + ((size_t)cta * 2 + ((cu0 >= tu0) ? 0 : 1)) * 64;
            mine[lane] = v;
            mine[32 + lane] = vu;
            __syncwarp();
            unsigned int ticket = 0;
            if (lane == 0) ticket = atom_add_acq_rel(&p.tickets[tile_id], 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket != (unsigned)(c_last - c_first)) continue;  // not the last CTA
            v = 0.f;
            vu = 0.f;
            for (int c = c_first; c <= c_last; c++) {
                const float* theirs = p.part + ((size_t)c * 2 + ((cta_first(c) >= tu0) ? 0 : 1)) * 64;
                v += ld_relaxed_gpu(theirs + lane);
                vu += ld_relaxed_gpu(theirs + 32 + lane);
            }
            if (lane == 0) p.tickets[tile_id] = 0;
        }
        epilogue(s, tile, v, vu);
    }
    MMA_STAMP(6);
}

// Pull the first ring stages of launch/phase `p` towards L2 (fire-and-forget): called by the megakernel for the
// NEXT phase before it waits at a grid barrier, so the weight stream does not stop at the phase boundary.
__device__ __forceinline__ void mma_warm_l2(const MParams& p, int depth) {
    if (p.expert_sel) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nw = blockDim.x >> 5;
    const int cta = blockIdx.x;
    if (cta >= p.n_ctas) return;
    const int first = cta * p.cbase + min(cta, p.crem), cnt = p.cbase + (cta < p.crem ? 1 : 0);
    const int cu0 = p.tile_mode ? first * p.units_per_tile : first;
    const int cn = p.tile_mode ? cnt * p.units_per_tile : cnt;
    const int u0 = cu0 + (warp * cn) / nw, u1 = cu0 + ((warp + 1) * cn) / nw;
    if (u1 == u0) return;
    const bool swiglu = p.epi == ME_SWIGLU;
    int s = (p.n_seg > 2 && u0 >= p.seg[2].unit0) ? 2 : (p.n_seg > 1 && !swiglu && u0 >= p.seg[1].unit0) ? 1 : 0;
    const int local = u0 - p.seg[s].unit0;
    const int tile = local / p.units_per_tile;
    int chunk = local - tile * p.units_per_tile, mat = s;
    if (swiglu && chunk >= p.chunks) { mat = 1; chunk -= p.chunks; }
    const MSeg& sg = p.seg[mat];
    // the run of this warp's first units inside its first row: up to `depth` units, contiguous bytes per row
    const int n = min(min(depth, u1 - u0), p.chunks - chunk);
    const long long off = (long long)chunk * sg.chunk_bytes;
    const int bytes = n * sg.chunk_bytes;
    for (int j = 0; j < 4; j++) {
        const uint8_t* a = sg.w + (long long)min(tile * kMmaRows + g + 8 * j, sg.n_rows - 1) * sg.row_bytes + off;
        for (int o = 128 * t; o < bytes + 127; o += 512) prefetch_l2(a + o);
    }
}

__global__ void __launch_bounds__(kMmaMaxWarps * 32, 1) gemv_mma_kernel(const __grid_constant__ MParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_red[2 * kMmaMaxWarps];
    __shared__ float s_part[kMmaMaxWarps][2][2][32];   // pieces of tiles shared between warps of this CTA
    mma_gemv_cta(p, smem, s_red, s_part, [] {}, [](auto&&) { pdl_launch_dependents(); pdl_wait(); }, false, true);
}

// ---------------------------------------------------------------- host-side launch planning
inline bool mma_type_ok(int type) { return type == T_Q4_K || type == T_Q5_K || type == T_Q6_K || type == T_Q8_0; }

// bytes between row slots: the unit's bytes plus up to 15 bytes of source misalignment, residue mod 128 chosen
// for conflict-free fragment loads (LDS.128 row pairs: 64; 32-bit loads of 8 rows: odd multiple of 16)
inline int mma_row_stride(int type) {
    const int cb = kMmaChunk / type_block_elems(type), bb = type_block_bytes(type);
    const bool aligned = (type == T_Q4_K || type == T_Q5_K);
    int rs = aligned ? ((cb * bb + 15) & ~15) : ((15 + cb * bb + 4 + 15) & ~15);
    for (;; rs += 16) {
        const int m = rs & 127;
        if (aligned ? (m == 64) : ((m & 15) == 0 && ((m >> 4) & 1))) return rs;
    }
}

// How the units of a launch are dealt to `grid` CTAs.  Whole tiles per CTA when that costs less than the ~2 us a
// merge through global memory takes (imbalance of one tile on a short launch); unit-balanced stream-K otherwise.
inline void mma_deal(MParams& p, int grid) {
    int tiles = 0;
    double bytes = 0;
    const int nseg_tiles = (p.epi == ME_SWIGLU) ? 1 : p.n_seg;
    for (int s = 0; s < nseg_tiles; s++) tiles += p.seg[s].n_tiles;
    for (int s = 0; s < p.n_seg; s++) bytes += (double)p.seg[s].n_rows * p.seg[s].nb_row * p.seg[s].bb;
    p.n_ctas = grid;
    p.tile_mode = 0;
    if (tiles >= grid) {
        const int per = (tiles + grid - 1) / grid;
        const double imbalance = (double)per * grid / tiles - 1.0;        // extra time of the fullest CTA
        const double cost_us = imbalance * bytes / 5.0e12 * 1e6;          // at ~5 TB/s
        if (cost_us < 1.5) p.tile_mode = 1;
    }
    const int n = p.tile_mode ? tiles : p.total_units;
    if (n < grid) p.n_ctas = n;
    p.cbase = n / p.n_ctas;
    p.crem = n % p.n_ctas;
}

struct MPlan {
    int grid, warps, stages;
    size_t smem;
};

// Fills the derived fields of p (segments' w/out/bias/row_bytes/expert_stride/type/n_rows, n_seg, K, epi
// must be set) and picks the ring depth so that x + the rings fit in shared memory.  Returns false if the
// launch is not eligible for this kernel (caller falls back to the CUDA-core kernel).
inline bool mma_plan(MParams& p, int n_sm, int want_warps, int want_stages, size_t smem_limit, MPlan& plan) {
    (void)want_warps;   // the block is always kMmaMaxWarps warps (the megakernel's block size)
    if (p.K % 32) return false;
    if (p.epi == ME_SWIGLU && (p.n_seg != 2 || p.seg[0].n_rows != p.seg[1].n_rows)) return false;
    int max_rs = 0;
    for (int s = 0; s < p.n_seg; s++) {
        MSeg& sg = p.seg[s];
        if (!mma_type_ok(sg.type)) return false;
        if (p.K % type_block_elems(sg.type)) return false;
        if ((sg.type == T_Q4_K || sg.type == T_Q5_K) && ((sg.row_bytes & 15) || (sg.expert_stride & 15) || ((uintptr_t)sg.w & 15))) return false;
        if ((sg.row_bytes & 1) || (sg.expert_stride & 1) || ((uintptr_t)sg.w & 1)) return false;
        sg.row_stride = mma_row_stride(sg.type);
        sg.cb = kMmaChunk / type_block_elems(sg.type);
        sg.bb = type_block_bytes(sg.type);
        sg.chunk_bytes = sg.cb * sg.bb;
        sg.nb_row = p.K / type_block_elems(sg.type);
        sg.n_tiles = (sg.n_rows + kMmaRows - 1) / kMmaRows;
        max_rs = std::max(max_rs, sg.row_stride);
    }
    p.chunks = (p.K + kMmaChunk - 1) / kMmaChunk;
    if (p.epi == ME_SWIGLU) {
        p.units_per_tile = 2 * p.chunks;
        p.seg[0].unit0 = 0;
        p.seg[1].unit0 = 0;
        p.total_units = p.seg[0].n_tiles * p.units_per_tile;
    } else {
        p.units_per_tile = p.chunks;
        int u = 0;
        for (int s = 0; s < p.n_seg; s++) {
            p.seg[s].unit0 = u;
            u += p.seg[s].n_tiles * p.units_per_tile;
        }
        p.total_units = u;
    }
    p.stage_bytes = kMmaRows * max_rs;
    const size_t xb = x_smem_bytes(p.K);
    const int warps = kMmaMaxWarps;
    int stages = std::max(2, std::min(want_stages, kMmaMaxStages));
    auto need = [&](int w, int st) { return xb + (size_t)w * st * p.stage_bytes + 16; };  // +16: funnel loads read one word past a piece
    while (need(warps, stages) > smem_limit) {
        if (stages > 2) stages--;
        else return false;
    }
    p.stages = stages;
    plan.warps = warps;
    plan.stages = stages;
    plan.smem = need(warps, stages);
    plan.grid = (int)std::min<long long>(n_sm, (p.total_units + warps - 1) / warps);
    if (plan.grid < 1) plan.grid = 1;
    mma_deal(p, plan.grid);
    return true;
}

using MmaKernel = void (*)(const MParams);
inline MmaKernel mma_kernel_for(int) { return gemv_mma_kernel; }
inline cudaError_t mma_set_smem_limit(int bytes) {
    return cudaFuncSetAttribute(gemv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace b200
