// gemv_mma.cuh — the production dequant-GEMV for batch-1 decode (Q4_K / Q5_K / Q6_K / Q8_0).
//
// y[j] = sum_k deq(W)[j,k] * x[k] on the untouched GGUF super-block layout (row j of W is
// K/bs contiguous blocks), replacing the reference's vec_mat_q* kernels
// (src/backend/cuda/kernels.rs:443-735) and the CPU hot loop fused_vecmat_dispatch ->
// simd::dot_q* (src/backend/cpu/ops.rs:1123-1191, src/backend/cpu/simd.rs:931-1146).
// Fused around it: RMSNorm of x (simd.rs:847-899), +bias (layers.rs:68-74), +residual
// (layers.rs:1202-1241), silu(gate)*up (simd.rs:598-649), out += w_e*y (moe.rs:363-368).
//
// Why this shape (profiles/r01_v1_*: the CUDA-core version spent ~6 issue slots per weight
// and stalled at 1.2 TB/s):
//  * weights go HBM -> shared memory with 16-byte cp.async (SASS LDGSTS) into PER-WARP rings.
//    A unit is 16 rows x 256 elements (one K-quant super-block per row); a lane moves 16-byte pieces of two rows, so an
//    instruction covers 8 rows x 64 contiguous bytes and costs no address arithmetic.  (Per-row
//    cp.async.bulk copies measured slower: UBLKCP is issued one lane at a time, ~10 issue slots
//    per 288-byte copy, profiles/r01_v2_*.)  A warp produces and consumes its own ring, so
//    there is no cross-warp hand-off and the first stages are issued BEFORE
//    griddepcontrol.wait (weights never depend on the previous kernel of the token);
//  * dot products run on the tensor pipe (mma.sync.m16n8k16, SASS HMMA): quants become exact
//    fp16 integers with one LOP3/PRMT per two elements ((w & 0x000F000F) | 0x6400_6400 =
//    1024+q), x is split into fp16 hi + lo parts (x = hi + lo to 2^-22) that sit in two
//    columns of the B operand, accumulation is f32.  Block scales / mins are applied in f32
//    to per-sub-block sums (the reference's separated form, simd.rs:1006-1013); the integer
//    bias (1024, +32 for Q6_K, +128 for Q8_0) is removed with per-16-element sums of x.  The
//    8 columns of the MMA are used as 4 (hi, lo) pairs: the B operand is zero except in the
//    pair of the sub-block a lane's k-slots belong to, so each sub-block's sum lands in the
//    lane that decoded its scale and nothing is shuffled until a tile is finished;
//  * stream-K: all (tile, chunk) units of a launch are dealt evenly, in contiguous runs, to the
//    warps of a persistent grid (one CTA per SM, up to 16 warps).  Tiles cut across warps are
//    merged through shared memory inside a CTA and through a small global scratch + ticket
//    across CTAs, always in a fixed order (run-to-run deterministic).
//
// ~1.3 issue slots per weight (Q4_K) instead of ~6; see DESIGN.md for the budget.
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

constexpr int kMmaMaxWarps = 16;
constexpr int kMmaMaxStages = 4;

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
    int n_tiles;           // ceil(n_rows / 16)
    int unit0;             // first unit of this segment in the launch
    int row_stride;        // pitch of the row slots of a ring stage
    int cb;                // blocks per unit
    int bb;                // block bytes
    int chunk_bytes;       // cb * bb
    int nb_row;            // blocks per row (K / block elems)
};

struct MParams {
    MSeg seg[3];
    int n_seg;
    int K;
    int chunks;            // ceil(K / 256)
    int units_per_tile;    // chunks (2*chunks for ME_SWIGLU: gate chunks then up chunks)
    int total_units;
    // dealing: CTA b < n_ctas owns `cbase` (+1 for b < crem) consecutive units (tile_mode 0) or whole 16-row tiles
    // (tile_mode 1: no tile straddles two CTAs, so nothing is merged through global memory); its warps split the
    // CTA's units evenly
    int n_ctas, tile_mode, cbase, crem;
    int stages;
    int stage_bytes;       // 16 * max row_stride
    const float* x;        // [K] f32
    const float* norm_w;   // optional fused RMSNorm weight [K]
    float eps;
    int epi;
    const float* residual;
    // MoE (expert-resident): the slot-th selected expert
    const int* expert_sel;
    const float* expert_wt;
    int expert_slot;
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
    // cross-CTA merge scratch
    float* part;             // [grid][2][32]
    unsigned int* tickets;   // [total logical tiles], zero between launches
    int* err;                // device error flag (unused by the cp.async pipeline; kept for the watchdog ABI)
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
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & mask) | orv
    return r;
}
// Shared-memory loads are NOT volatile so that ptxas/nvcc may interleave them with the MMAs of neighbouring
// blocks.  Ordering against the cp.async pipeline comes from data dependencies: every address is derived from an
// opaque token produced after the wait (smem_token), and the unit's results are pinned before the stage is refilled.
__device__ __forceinline__ uint32_t smem_token() {
    uint32_t t;
    asm volatile("mov.u32 %0, 0;" : "=r"(t)::"memory");
    return t;
}
__device__ __forceinline__ void pin2(float& a, float& b) { asm volatile("" : "+f"(a), "+f"(b)::"memory"); }
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
// 16 bytes at a 2-byte-aligned shared address: 5 aligned words + funnel shifts (shift 0 or 16)
__device__ __forceinline__ void lds_piece16(uint32_t a, uint32_t (&o)[4]) {
    const uint32_t base = a & ~3u, sh = (a & 3u) << 3;
    const uint32_t w0 = lds32(base), w1 = lds32(base + 4), w2 = lds32(base + 8), w3 = lds32(base + 12), w4 = lds32(base + 16);
    o[0] = __funnelshift_r(w0, w1, sh);
    o[1] = __funnelshift_r(w1, w2, sh);
    o[2] = __funnelshift_r(w2, w3, sh);
    o[3] = __funnelshift_r(w3, w4, sh);
}
// 4 bytes at a 2-byte-aligned shared address
__device__ __forceinline__ uint32_t lds32_a2(uint32_t a) {
    const uint32_t base = a & ~3u, sh = (a & 3u) << 3;
    return __funnelshift_r(lds32(base), lds32(base + 4), sh);
}

// Quants as fp16 operands: the integer q sits in the low mantissa bits of an fp16 with a zero exponent field,
// i.e. it IS the subnormal q * 2^-24 -- no magic-number offset (the classic 0x6400 | q = 1024 + q form makes the
// f32 accumulator carry 1024 * sum(x) and costs ~7 bits: measured 3e-5 vs 1e-6 relative error).  The 2^24 is
// folded into the final scale and into the staged sums of x ("operand units").
constexpr uint32_t kMagic = 0u;
constexpr uint32_t kMagicB = 0u;                      // PRMT filler byte
constexpr float kXsScale = 5.9604644775390625e-08f;   // 2^-24
constexpr float kUnscale = 16777216.0f;               // 2^24

// ---------------------------------------------------------------- x in shared memory
// fp16 hi and lo parts (x ~= hi + lo), stored so that the 4 elements starting at e (e % 4 == 0)
// read as one 8-byte word give the m16n8k16 B fragment of a lane whose four k-slots come from one
// 32-bit word of quants: order [x0, x2, x1, x3] (bytes 0,2 -> k-slots 2t,2t+1; bytes 1,3 -> 2t+8,2t+9).
// xs16[i] = sum of float(hi)+float(lo) over elements 16i..16i+15.
struct XSmem {
    uint32_t xh, xl, xs;  // shared-space byte addresses: hi halves, lo halves, per-16 sums
    uint32_t s32;         // per-32-element scale 2^-k of the staged values
    uint32_t zero;        // 256 bytes of zeros: the B operand of lanes whose column pair is not addressed
};
__host__ __device__ __forceinline__ int xperm(int e) { return (e & ~3) | (((e & 1) << 1) | ((e >> 1) & 1)); }
// the lo array sits 64 bytes off a multiple of 128 from the hi array: the hi and lo lanes of a B-fragment
// load hit different banks
constexpr uint32_t kXlPad = 64;
// hi[K] + pad + lo[K] halves, xs16[K/16] floats, s32[K/32] floats, 256 zero bytes
__host__ __device__ inline size_t x_smem_bytes(int K) {
    return (((size_t)4 * K + kXlPad + (size_t)(K >> 2) + (size_t)(K >> 3) + 127) & ~(size_t)127) + 256;
}

// x staging, all threads of the CTA, K % 32 == 0, ONE pass and one __syncthreads (by the caller):
//   y = x * w (w: optional RMSNorm weight; the scalar 1/rms is applied to the finished dot products instead of
//   to every element -- (x*inv)*w in the reference, simd.rs:891-892, differs by one rounding);
//   every 32 elements are scaled by their own power of two 2^k (exact) so that the largest sits near 2^12, then
//   split into fp16 hi + lo: ~22 significant bits relative to the largest element of the group, whatever |x|;
//   xs16 = sums of the true values (in operand units, x 2^-24), s32 = 2^-k, red[warp] = partial sum of x^2.
// stage_x_load issues the global loads (first kXRegs float4 per thread) BEFORE the caller issues weight copies.
constexpr int kXRegs = 4;
struct XStage {
    float4 v[kXRegs], w[kXRegs];
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
__device__ __forceinline__ void stage_x_load(XStage& st, const XSource& xs, const float* __restrict__ norm_w, int K) {
    const int tid = threadIdx.x, nthr = blockDim.x;
#pragma unroll
    for (int i = 0; i < kXRegs; i++) {
        const int e = (tid + i * nthr) * 4;
        st.v[i] = (e < K) ? x_fetch4(xs, e) : make_float4(0.f, 0.f, 0.f, 0.f);
        st.w[i] = (norm_w && e < K) ? *reinterpret_cast<const float4*>(norm_w + e) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
}
__device__ __forceinline__ float split_store4(float4 v, float4 w, int e, __half* xh, __half* xl, float* xs, float* s32, unsigned mask) {
    const float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
    float am = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    am = fmaxf(am, __shfl_xor_sync(mask, am, 1));
    am = fmaxf(am, __shfl_xor_sync(mask, am, 2));
    am = fmaxf(am, __shfl_xor_sync(mask, am, 4));
    int k = 0;
    if (am > 0.0f && am < 3.0e38f) k = min(max(138 - (int)((__float_as_uint(am) >> 23) & 0xFFu), -100), 100);  // 12 - (exp - 126)
    const float up = __int_as_float((127 + k) << 23), down = __int_as_float((127 - k) << 23);
    v.x *= up; v.y *= up; v.z *= up; v.w *= up;
    const __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z), h3 = __float2half_rn(v.w);
    const __half l0 = __float2half_rn(v.x - __half2float(h0)), l1 = __float2half_rn(v.y - __half2float(h1));
    const __half l2 = __float2half_rn(v.z - __half2float(h2)), l3 = __float2half_rn(v.w - __half2float(h3));
    uint2 ph, pl;  // stored order [x0, x2, x1, x3]
    ph.x = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h2) << 16);
    ph.y = (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h3) << 16);
    pl.x = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l2) << 16);
    pl.y = (uint32_t)__half_as_ushort(l1) | ((uint32_t)__half_as_ushort(l3) << 16);
    *reinterpret_cast<uint2*>(xh + e) = ph;
    *reinterpret_cast<uint2*>(xl + e) = pl;
    float sum = ((__half2float(h0) + __half2float(l0)) + (__half2float(h1) + __half2float(l1))) +
                ((__half2float(h2) + __half2float(l2)) + (__half2float(h3) + __half2float(l3)));
    sum += __shfl_xor_sync(mask, sum, 1);
    sum += __shfl_xor_sync(mask, sum, 2);
    if ((threadIdx.x & 3) == 0) xs[e >> 4] = sum * (down * kXsScale);
    if ((threadIdx.x & 7) == 0) s32[e >> 5] = down;
    return ss;
}
__device__ __forceinline__ void stage_x_finish(const XStage& st, const XSource& xsrc, const float* __restrict__ norm_w, int K,
                                               uint8_t* smem, float* red /*[kMmaMaxWarps]*/) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __half* xh = reinterpret_cast<__half*>(smem);
    __half* xl = reinterpret_cast<__half*>(smem + (size_t)2 * K + kXlPad);
    float* xs = reinterpret_cast<float*>(smem + (size_t)4 * K + kXlPad);
    float* s32 = xs + (K >> 4);
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < kXRegs; i++) {
        const int e = (tid + i * nthr) * 4;
        if (e < K) ss += split_store4(st.v[i], st.w[i], e, xh, xl, xs, s32, __activemask());
    }
    for (int e = (tid + kXRegs * nthr) * 4; e < K; e += nthr * 4) {
        const float4 v = x_fetch4(xsrc, e);
        const float4 w = norm_w ? *reinterpret_cast<const float4*>(norm_w + e) : make_float4(1.f, 1.f, 1.f, 1.f);
        ss += split_store4(v, w, e, xh, xl, xs, s32, __activemask());
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    if (tid < 64) reinterpret_cast<uint32_t*>(smem + (x_smem_bytes(K) - 256))[tid] = 0u;
}
// after the caller's __syncthreads: the factor for finished dot products (1/rms when normalising, and 2^24)
__device__ __forceinline__ float stage_x_unscale(const float* red, bool norm, float eps, int K) {
    float inv = 1.0f;
    if (norm) {
        float tot = 0.0f;
        const int nwarp = blockDim.x >> 5;
        for (int w = 0; w < nwarp; w++) tot += red[w];
        inv = 1.0f / sqrtf(tot / (float)K + eps);
    }
    return inv * kUnscale;
}

// ---------------------------------------------------------------- per-type unit kernels
// sp = shared address of the stage (row slot r at sp + r*RS; the row's bytes start at +doff),
// nblk blocks, e0 = element index of the chunk's first element, g = lane>>2 (rows g and g+8),
// t = lane&3.  They add this unit's contribution for rows g and g+8 to acc0 / acc1 (partial over
// t: summed when the tile is finished).

// get_scale_min_k4 (dequant.rs:213-225) for sub-blocks 2t (l) and 2t+1 (h) of a block header h = {d|dmin, scales[12]}.
// One formula for both halves of the table: sub-blocks 0..3 are 6-bit fields of bytes 0..7, sub-blocks 4..7 take
// their low 4 bits from bytes 8..11 and their top 2 bits from bits 6,7 of bytes 0..7.
struct K4Lane {
    uint32_t sh_lo, sh_w, sh_w4, m_lo, m_w;
};
__device__ __forceinline__ K4Lane k4_lane(int t) {
    K4Lane k;
    const uint32_t sh = 16u * (uint32_t)(t & 1);
    k.sh_lo = (t < 2) ? sh : sh + 2u;
    k.m_lo = (t < 2) ? 0x3F3Fu : 0x3030u;
    k.m_w = (t < 2) ? 0u : 0x0F0Fu;
    k.sh_w = sh;
    k.sh_w4 = sh + 4u;
    return k;
}
__device__ __forceinline__ void k4_scales(const uint4& h, const K4Lane& k, float& dl, float& ml, float& dh, float& mh) {
    const float d = half_bits_to_float(h.x), dmin = half_bits_to_float(h.x >> 16);
    const uint32_t scp = lop3_and_or(h.y >> k.sh_lo, k.m_lo, (h.w >> k.sh_w) & k.m_w);
    const uint32_t mnp = lop3_and_or(h.z >> k.sh_lo, k.m_lo, (h.w >> k.sh_w4) & k.m_w);
    dl = d * (float)(scp & 0xFFu);
    dh = d * (float)((scp >> 8) & 0xFFu);
    ml = dmin * (float)(mnp & 0xFFu);
    mh = dmin * (float)((mnp >> 8) & 0xFFu);
}

// Q4_K / Q5_K (blocks.rs:114-141).  Lane t reads 16 qs bytes per 64-byte half c of the block:
// bytes 64c+16t.. -> group gp = 2c + (t>>1) (64 elements: low nibbles = sub-block 2gp, high = 2gp+1),
// positions l = 16(t&1) + 4i + j.  Group gp is routed to column pair gp, so lane t' of the D
// fragment ends up with sub-blocks 2t' (cl) and 2t'+1 (ch), whose scales it decodes.
template <bool Q5>
__device__ __forceinline__ void unit_k45(uint32_t sp, uint32_t RS, int nblk, int e0, const XSmem& sm, int g, int t, float& acc0,
                                         float& acc1) {
    constexpr int BB = Q5 ? 176 : 144, QS = Q5 ? 48 : 16;
    const uint32_t arr = (g & 1) ? sm.xl : sm.xh;
    const bool lane_act = (((g >> 1) & 1) == (t >> 1));
    const int c_act = g >> 2;
    const K4Lane kl = k4_lane(t);
    // B-operand base of this lane for each half c: its x elements, or the zero page when its column pair is not addressed
    const uint32_t xo = 2u * (uint32_t)(e0 + 128 * c_act + 64 * (t >> 1) + 16 * (t & 1));
    const uint32_t xb0 = (lane_act && c_act == 0) ? arr + xo : sm.zero, xs0 = (lane_act && c_act == 0) ? 512u : 0u;
    const uint32_t xb1 = (lane_act && c_act == 1) ? arr + xo : sm.zero, xs1 = (lane_act && c_act == 1) ? 512u : 0u;
#pragma unroll 2
    for (int b = 0; b < nblk; b++) {
        const uint32_t r0 = sp + g * RS + b * BB, r1 = r0 + 8 * RS;
        const uint4 h0 = lds128(r0), h1 = lds128(r1);
        uint4 qh0 = make_uint4(0u, 0u, 0u, 0u), qh1 = qh0;
        if (Q5) {
            qh0 = lds128(r0 + 16 + 16 * (t & 1));
            qh1 = lds128(r1 + 16 + 16 * (t & 1));
        }
        // independent accumulator chains per half c (4 MMAs deep instead of 8): cl[c] / ch[c]
        float cl[2][4], ch[2][4];
#pragma unroll
        for (int c = 0; c < 2; c++) cl[c][0] = cl[c][1] = cl[c][2] = cl[c][3] = ch[c][0] = ch[c][1] = ch[c][2] = ch[c][3] = 0.f;
        const int eb = e0 + b * 256;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint4 W0 = lds128(r0 + QS + 64 * c + 16 * t), W1 = lds128(r1 + QS + 64 * c + 16 * t);
            const int gp = 2 * c + (t >> 1);
            const uint32_t xa = (c ? xb1 : xb0) + (uint32_t)b * (c ? xs1 : xs0);
            uint4 bl[2], bh[2];
            bl[0] = lds128(xa);
            bl[1] = lds128(xa + 16);
            bh[0] = lds128(xa + 64);
            bh[1] = lds128(xa + 80);
            const uint32_t wa4[4] = {W0.x, W0.y, W0.z, W0.w}, wb4[4] = {W1.x, W1.y, W1.z, W1.w};
            const uint32_t ha4[4] = {qh0.x, qh0.y, qh0.z, qh0.w}, hb4[4] = {qh1.x, qh1.y, qh1.z, qh1.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t wa = wa4[i], wb = wb4[i];
                const uint32_t blx = (i & 1) ? bl[i >> 1].z : bl[i >> 1].x, bly = (i & 1) ? bl[i >> 1].w : bl[i >> 1].y;
                const uint32_t bhx = (i & 1) ? bh[i >> 1].z : bh[i >> 1].x, bhy = (i & 1) ? bh[i >> 1].w : bh[i >> 1].y;
                uint32_t ml_a = kMagic, ml_b = kMagic, ml_a8 = kMagic, ml_b8 = kMagic;  // low-group or-values
                uint32_t mh_a = kMagic, mh_b = kMagic, mh_a8 = kMagic, mh_b8 = kMagic;  // high-group or-values
                if (Q5) {  // 5th bit (dequant.rs:262-315): +16 for the low sub-block (bit 4), +256 (= 16*16) for the x16-carried high one
                    const uint32_t la = ha4[i] >> (2 * gp), lb = hb4[i] >> (2 * gp);  // bit0 of each byte: sub-block 2gp, bit1: 2gp+1
                    ml_a = lop3_and_or(la << 4, 0x00100010u, kMagic);
                    ml_b = lop3_and_or(lb << 4, 0x00100010u, kMagic);
                    ml_a8 = lop3_and_or(la >> 4, 0x00100010u, kMagic);
                    ml_b8 = lop3_and_or(lb >> 4, 0x00100010u, kMagic);
                    mh_a = lop3_and_or(la << 7, 0x01000100u, kMagic);
                    mh_b = lop3_and_or(lb << 7, 0x01000100u, kMagic);
                    mh_a8 = lop3_and_or(la >> 1, 0x01000100u, kMagic);
                    mh_b8 = lop3_and_or(lb >> 1, 0x01000100u, kMagic);
                }
                const uint32_t wa8 = __umulhi(wa, 0x01000000u), wb8 = __umulhi(wb, 0x01000000u);  // >> 8 on the FMA pipe
                mma16816(cl[c], lop3_and_or(wa, 0x000F000Fu, ml_a), lop3_and_or(wb, 0x000F000Fu, ml_b),
                         lop3_and_or(wa8, 0x000F000Fu, ml_a8), lop3_and_or(wb8, 0x000F000Fu, ml_b8), blx, bly);
                mma16816(ch[c], lop3_and_or(wa, 0x00F000F0u, mh_a), lop3_and_or(wb, 0x00F000F0u, mh_b),
                         lop3_and_or(wa8, 0x00F000F0u, mh_a8), lop3_and_or(wb8, 0x00F000F0u, mh_b8), bhx, bhy);
            }
        }
        // lane t owns sub-blocks 2t (low nibbles) and 2t+1 (high nibbles, carried x16) of this block: column pair t,
        // which was fed by the half c = t>>1
        const float sl0 = (t & 2) ? cl[1][0] + cl[1][1] : cl[0][0] + cl[0][1], sl1 = (t & 2) ? cl[1][2] + cl[1][3] : cl[0][2] + cl[0][3];
        const float sh0 = (t & 2) ? ch[1][0] + ch[1][1] : ch[0][0] + ch[0][1], sh1 = (t & 2) ? ch[1][2] + ch[1][3] : ch[0][2] + ch[0][3];
        float dl0, ml0, dh0, mh0, dl1, ml1, dh1, mh1;
        k4_scales(h0, kl, dl0, ml0, dh0, mh0);
        k4_scales(h1, kl, dl1, ml1, dh1, mh1);
        const uint32_t xsa = sm.xs + 4u * (uint32_t)((eb >> 4) + 4 * t);
        const float xsl = lds_f32(xsa) + lds_f32(xsa + 4), xsh = lds_f32(xsa + 8) + lds_f32(xsa + 12);
        const uint2 sc = lds64(sm.s32 + 4u * (uint32_t)((eb >> 5) + 2 * t));  // 2^-k of sub-blocks 2t, 2t+1
        const float kl = __uint_as_float(sc.x), kh = __uint_as_float(sc.y) * 0.0625f;
        acc0 += (dl0 * kl) * sl0 - ml0 * xsl + (dh0 * kh) * sh0 - mh0 * xsh;
        acc1 += (dl1 * kl) * sl1 - ml1 * xsl + (dh1 * kh) * sh1 - mh1 * xsh;
    }
}

// Q6_K (blocks.rs:143-155, dequant.rs:321-356): ql[128] qh[64] scales[16] d.  16 scale groups of 16.
// Per 128-half n, lane t reads ql bytes 64n+16t.. : t<2 -> "A" bytes (low nibble: quarter c=0, high: c=2),
// t>=2 -> "B" bytes (c=1 / c=3); positions l = 16(t&1)+4i+j, so lane t's low-nibble values are elements
// 128n + 16t + 4i + j and its high-nibble values are +64.  Lane t's group (scale 8n + 2c + (t&1) = 4m + t,
// m = 2n + lowhigh) is routed to column pair t: four accumulator sets C[m], D-lane t' owns scale 4m + t'.
__device__ __forceinline__ void unit_q6k(uint32_t sp, uint32_t RS, int nblk, int e0, uint32_t doff0, uint32_t doff1,
                                         const XSmem& sm, int g, int t, float& acc0, float& acc1) {
    const uint32_t arr = (g & 1) ? sm.xl : sm.xh;
    const bool act = (g >> 1) == t;
    const uint32_t s_lo = 4u - 2u * (uint32_t)(t >> 1);  // (qh >> 2c) << 4 for c = t>>1
    const uint32_t s_hi = 2u * (uint32_t)(t >> 1);       // (qh >> 2c) << 4 for c = 2 + (t>>1): qh >> (2*(t>>1)), bits 4,5
    for (int b = 0; b < nblk; b++) {
        const uint32_t r0 = sp + g * RS + doff0 + b * 210, r1 = sp + (g + 8) * RS + doff1 + b * 210;
        float C[4][4];
#pragma unroll
        for (int m = 0; m < 4; m++) C[m][0] = C[m][1] = C[m][2] = C[m][3] = 0.f;
        const int eb = e0 + b * 256;
#pragma unroll
        for (int n = 0; n < 2; n++) {
            uint32_t L0[4], L1[4], H0[4], H1[4];
            lds_piece16(r0 + 64 * n + 16 * t, L0);
            lds_piece16(r1 + 64 * n + 16 * t, L1);
            lds_piece16(r0 + 128 + 32 * n + 16 * (t & 1), H0);
            lds_piece16(r1 + 128 + 32 * n + 16 * (t & 1), H1);
            const uint32_t xa = act ? arr + 2u * (uint32_t)(eb + 128 * n + 16 * t) : sm.zero;
            uint4 bl[2], bh[2];
            bl[0] = lds128(xa);
            bl[1] = lds128(xa + 16);
            bh[0] = lds128(xa + 128);
            bh[1] = lds128(xa + 144);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t lo0 = lop3_and_or(L0[i], 0x0F0F0F0Fu, (H0[i] << s_lo) & 0x30303030u);
                const uint32_t lo1 = lop3_and_or(L1[i], 0x0F0F0F0Fu, (H1[i] << s_lo) & 0x30303030u);
                const uint32_t hi0 = lop3_and_or(L0[i] >> 4, 0x0F0F0F0Fu, (H0[i] >> s_hi) & 0x30303030u);
                const uint32_t hi1 = lop3_and_or(L1[i] >> 4, 0x0F0F0F0Fu, (H1[i] >> s_hi) & 0x30303030u);
                const uint32_t blx = (i & 1) ? bl[i >> 1].z : bl[i >> 1].x, bly = (i & 1) ? bl[i >> 1].w : bl[i >> 1].y;
                const uint32_t bhx = (i & 1) ? bh[i >> 1].z : bh[i >> 1].x, bhy = (i & 1) ? bh[i >> 1].w : bh[i >> 1].y;
                mma16816(C[2 * n], __byte_perm(lo0, kMagicB, 0x4240), __byte_perm(lo1, kMagicB, 0x4240),
                         __byte_perm(lo0, kMagicB, 0x4341), __byte_perm(lo1, kMagicB, 0x4341), blx, bly);
                mma16816(C[2 * n + 1], __byte_perm(hi0, kMagicB, 0x4240), __byte_perm(hi1, kMagicB, 0x4240),
                         __byte_perm(hi0, kMagicB, 0x4341), __byte_perm(hi1, kMagicB, 0x4341), bhx, bhy);
            }
        }
        const float d0 = half_bits_to_float(lds16(r0 + 208)), d1 = half_bits_to_float(lds16(r1 + 208));
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int si = 4 * m + t;
            const float xs = lds_f32(sm.xs + 4u * (uint32_t)((eb >> 4) + si));
            const float kk = lds_f32(sm.s32 + 4u * (uint32_t)((eb >> 5) + (si >> 1)));
            const float s0 = (float)lds_s8(r0 + 192 + si), s1 = (float)lds_s8(r1 + 192 + si);
            acc0 += (d0 * s0) * (kk * (C[m][0] + C[m][1]) - 32.0f * xs);   // Q6_K's -32
            acc1 += (d1 * s1) * (kk * (C[m][2] + C[m][3]) - 32.0f * xs);
        }
    }
}

// Q8_0 (blocks.rs:60-70): 34-byte blocks of 32.  Lane t reads bytes 16m+4t.. of a block (two MMAs per block);
// blocks 4i..4i+3 share one accumulator set through the column pairs (block bi -> pair bi), D-lane t' owns
// block 4i + t'.
__device__ __forceinline__ void unit_q80(uint32_t sp, uint32_t RS, int nblk, int e0, uint32_t doff0, uint32_t doff1,
                                         const XSmem& sm, int g, int t, float& acc0, float& acc1) {
    const uint32_t arr = (g & 1) ? sm.xl : sm.xh;
    const uint32_t row0 = sp + g * RS + doff0, row1 = sp + (g + 8) * RS + doff1;
    for (int b4 = 0; b4 < nblk; b4 += 4) {
        float C[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int bi = 0; bi < 4; bi++) {
            const int b = b4 + bi;
            if (b < nblk) {  // warp-uniform
                const uint32_t r0 = row0 + b * 34 + 2, r1 = row1 + b * 34 + 2;
                const bool act = (g >> 1) == bi;
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    const uint32_t w0 = lds32_a2(r0 + 16 * m + 4 * t) ^ 0x80808080u;  // int8 -> biased uint8
                    const uint32_t w1 = lds32_a2(r1 + 16 * m + 4 * t) ^ 0x80808080u;
                    const uint2 bf = lds64(act ? arr + 2u * (uint32_t)(e0 + 32 * b + 16 * m + 4 * t) : sm.zero);
                    mma16816(C, __byte_perm(w0, kMagicB, 0x4240), __byte_perm(w1, kMagicB, 0x4240),
                             __byte_perm(w0, kMagicB, 0x4341), __byte_perm(w1, kMagicB, 0x4341), bf.x, bf.y);
                }
            }
        }
        const int b = b4 + t;  // lane t owns block b4 + t
        if (b < nblk) {
            const uint32_t xa = sm.xs + 4u * (uint32_t)(((e0 + 32 * b) >> 4));
            const float xs = lds_f32(xa) + lds_f32(xa + 4);
            const float kk = lds_f32(sm.s32 + 4u * (uint32_t)((e0 >> 5) + b));
            acc0 += half_bits_to_float(lds16(row0 + b * 34)) * (kk * (C[0] + C[1]) - 128.0f * xs);  // int8 -> biased uint8
            acc1 += half_bits_to_float(lds16(row1 + b * 34)) * (kk * (C[2] + C[3]) - 128.0f * xs);
        }
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

// Position of a warp in its run of units: segment s, 16-row tile, matrix (ME_SWIGLU: 0 = gate, 1 = up; else = s),
// chunk (256 elements) within the row.  a / b are the true addresses of the unit's bytes in rows tile*16+g and +8.
struct MCursor {
    int s, tile, mat, chunk;
    const uint8_t* a;
    const uint8_t* b;
};

// The whole GEMV of one launch (or of one phase of the per-token megakernel, mega.cuh) for this CTA.
//   smem   : dynamic shared memory (x staging + rings), 128-byte aligned
//   s_red  : [2 * kMmaMaxWarps] floats, s_part: [kMmaMaxWarps][2][32] floats (static shared memory of the caller)
//   pdl    : the launch is part of a programmatic-dependent-launch chain (griddepcontrol at the right place)
//   warm_l2: pull the first stages towards L2 before anything that depends on the previous kernel
//   pre()  : runs first (megakernel: __syncthreads + arrive at the grid barrier that ends the previous phase)
//   post() : runs after the prologue, before anything that depends on other CTAs / the previous kernel
//            (megakernel: wait at that barrier; stand-alone kernel: griddepcontrol)
//   early  : issue the first ring stages with cp.async BEFORE post() (weights never depend on a predecessor), so
//            they land while the barrier is being waited for; otherwise they are issued after the x loads
template <int STAGES, class Pre, class Post>
__device__ __forceinline__ void mma_gemv_cta(const MParams& p, uint8_t* smem, float* s_red, float (*s_part)[2][32], Pre pre_fn,
                                             Post post_fn, bool early, bool warm_l2) {
    pre_fn();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nw = blockDim.x >> 5;
    const int K = p.K;
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
    long long eoff = 0;  // MoE expert index (valid after pdl_wait)
    MMA_STAMP(0);

    auto cur_ptrs = [&](MCursor& q) {
        const MSeg& sg = p.seg[q.mat];
        const uint8_t* base = sg.w + eoff * sg.expert_stride + (long long)q.chunk * sg.chunk_bytes;
        q.a = base + (long long)min(q.tile * 16 + g, sg.n_rows - 1) * sg.row_bytes;
        q.b = base + (long long)min(q.tile * 16 + g + 8, sg.n_rows - 1) * sg.row_bytes;
    };
    auto cur_init = [&](MCursor& q, int u) {
        q.s = (p.n_seg > 2 && u >= p.seg[2].unit0) ? 2 : (p.n_seg > 1 && !swiglu && u >= p.seg[1].unit0) ? 1 : 0;
        const int local = u - p.seg[q.s].unit0;
        q.tile = local / p.units_per_tile;
        q.chunk = local - q.tile * p.units_per_tile;
        q.mat = q.s;
        if (swiglu && q.chunk >= p.chunks) { q.mat = 1; q.chunk -= p.chunks; }
        cur_ptrs(q);
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
        cur_ptrs(q);
    };

    // ---- producer: cp.async the 16 rows of a unit into a ring stage.  Lane (g, t) moves the 16-byte pieces
    // t, t+4, ... of rows g and g+8 (8 rows x 64 contiguous bytes per instruction); sources are aligned down to
    // 16 bytes, the residue (doff) is re-derived by the consumer from the same address. ----
    MCursor cp{};
    int p_bytes_full = 0, p_bytes_last = 0, p_cbytes = 0;
    uint32_t p_dst = 0, p_rs8 = 0;   // this lane's first destination in stage 0, 8 row slots further
    auto prod_run = [&]() {  // per-run constants of the producer's matrix
        const MSeg& sg = p.seg[cp.mat];
        p_bytes_full = sg.chunk_bytes;
        p_bytes_last = (sg.nb_row - (p.chunks - 1) * sg.cb) * sg.bb;
        p_cbytes = sg.chunk_bytes;
        p_dst = ring + (uint32_t)g * sg.row_stride + 16u * t;
        p_rs8 = 8u * sg.row_stride;
    };
    auto issue = [&](uint32_t stage_off, bool l2_only) {
        const int bytes = (cp.chunk == p.chunks - 1) ? p_bytes_last : p_bytes_full;
        const uint32_t da = (uint32_t)((uintptr_t)cp.a & 15u), db = (uint32_t)((uintptr_t)cp.b & 15u);
        if (l2_only) {
            if (t * 128 < (int)da + bytes) prefetch_l2(cp.a - da + 128 * t);
            if (t * 128 < (int)db + bytes) prefetch_l2(cp.b - db + 128 * t);
        } else {
            const uint8_t* sa = cp.a - da + 16 * t;
            const uint8_t* sb = cp.b - db + 16 * t;
            const uint32_t dst = p_dst + stage_off, dstb = dst + p_rs8;
            const int ea = (int)da + bytes - 16 * t, eb = (int)db + bytes - 16 * t;  // bytes from this lane's first piece to the end
#pragma unroll
            for (int i = 0; i < 5; i++) {  // <= 287 bytes per row and unit
                if (64 * i < ea) cp_async16(dst + 64 * i, sa + 64 * i);
                if (64 * i < eb) cp_async16(dstb + 64 * i, sb + 64 * i);
            }
        }
        cp.chunk++;
        if (cp.chunk < p.chunks) {
            cp.a += p_cbytes;
            cp.b += p_cbytes;
        } else {
            cur_wrap(cp);
            prod_run();
        }
    };

    MCursor cc{};
    const int pre = min(STAGES - 1, n_units);
    if (n_units > 0 && !p.expert_sel) {
        // dense weights never depend on a predecessor: pull the first stages towards L2 before the PDL wait
        // (fire-and-forget; an early cp.async would make the x loads below queue behind DRAM-latency copies)
        cur_init(cp, u0);
        cc = cp;
        prod_run();
        if (early) {
#pragma unroll
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
#pragma unroll
        for (int k = 0; k < STAGES - 1; k++) cp_async_commit();
    }
    const bool issued_early = early && !p.expert_sel;

    post_fn();
    MMA_STAMP(1);

    if (p.expert_sel) {
        eoff = (long long)p.expert_sel[p.expert_slot];
        if (n_units > 0) { cur_init(cp, u0); cc = cp; prod_run(); }
    }
    // x: first pass (loads + sum of squares / max) is issued BEFORE the weight copies, the split after them
    XStage xst;
    const XSource xsrc{p.x, p.xsum, p.n_sum, p.sum_stride, p.x_res, p.x_full_out};
    stage_x_load(xst, xsrc, p.norm_w, K);
    if (!issued_early) {
#pragma unroll
        for (int k = 0; k < STAGES - 1; k++) {
            if (k < pre) issue((uint32_t)k * p.stage_bytes, false);
            cp_async_commit();
        }
    }
    MMA_STAMP(2);
    stage_x_finish(xst, xsrc, p.norm_w, K, smem, s_red);
    __syncthreads();
    const float unscale = stage_x_unscale(s_red, p.norm_w != nullptr, p.eps, K);
    MMA_STAMP(3);
    const uint32_t tokx = smem_token();
    XSmem sm;
    sm.xh = sbase + tokx;
    sm.xl = sm.xh + 2u * K + kXlPad;
    sm.xs = sm.xh + 4u * K + kXlPad;
    sm.s32 = sm.xs + (uint32_t)(K >> 2);
    sm.zero = sbase + tokx + (uint32_t)x_smem_bytes(K) - 256u;

    // ---- epilogue of a finished tile: lane L < 16 owns row tile*16 + L of segment s (v: lanes 16..31 = up rows) ----
    auto epilogue = [&](int s, int tile, float v) {
        const MSeg& sg = p.seg[s];
        const int j = tile * 16 + (lane & 15);
        const bool valid = (lane < 16) && (j < sg.n_rows);
        v *= unscale;
        float val = v;
        if (swiglu) {
            const float up = __shfl_sync(0xffffffffu, v, (lane & 15) + 16);
            val = mma_silu(v) * up;
        }
        if (valid) {
            if (sg.bias) val += sg.bias[j];
            if (p.epi == ME_RESIDUAL) val += p.residual[j];
            if (p.epi == ME_SCALED_ACC) {  // moe.rs:363-368
                const float prev = p.expert_slot == 0 ? 0.0f : sg.out[j];
                val = prev + p.expert_wt[p.expert_slot] * val;
                if (p.residual) val += p.residual[j];
            }
            if (p.n_peer > 0) {
                for (int r = 0; r < p.n_peer; r++) p.peer_out[r][j] = val;  // partial of a row-parallel GEMV, to every rank
            } else {
                sg.out[j] = val;
            }
        }
    };

    float ag0 = 0.f, ag1 = 0.f, au0 = 0.f, au1 = 0.f;
    int piece_s[2] = {-1, -1}, piece_tile[2] = {0, 0};  // tiles of which this warp holds only a piece (first / last of its run)
    uint32_t off_c = 0, off_p = (uint32_t)((STAGES - 1) % STAGES) * p.stage_bytes;  // consumer / producer stage offsets
    const uint32_t ring_bytes = (uint32_t)STAGES * p.stage_bytes;
    int issued = pre, done = 0;
    bool first = true;
    while (done < n_units) {
        // ---- a run: the consecutive units of this warp inside one row-tile of one matrix ----
        const int len = min(p.chunks - cc.chunk, n_units - done);
        const MSeg& wsg = p.seg[cc.mat];
        const int type = wsg.type;
        const uint32_t RS = (uint32_t)wsg.row_stride;
        const int cbytes = wsg.chunk_bytes;
        int e0 = cc.chunk * kMmaChunk;
        uint32_t ca = (uint32_t)(uintptr_t)cc.a, cb = (uint32_t)(uintptr_t)cc.b;  // low address bits: source misalignment
        float r0 = 0.f, r1 = 0.f;
        // one unit: keep the ring full, wait for the oldest stage, consume it
#define MMA_UNIT(CALL)                                                             \
    for (int i = 0; i < len; i++) {                                                \
        if (issued < n_units) { issue(off_p, false); issued++; }                   \
        off_p = (off_p + p.stage_bytes == ring_bytes) ? 0u : off_p + p.stage_bytes; \
        cp_async_commit();                                                         \
        cp_async_wait<STAGES - 1>();                                               \
        __syncwarp();                                                              \
        if (first) { MMA_STAMP(4); first = false; }                                \
        const uint32_t sp = ring + off_c + smem_token();                           \
        off_c = (off_c + p.stage_bytes == ring_bytes) ? 0u : off_c + p.stage_bytes; \
        float a0 = 0.f, a1 = 0.f;                                                  \
        CALL;                                                                      \
        pin2(a0, a1); /* the unit's shared-memory reads are complete before the stage can be refilled */ \
        r0 += a0;                                                                  \
        r1 += a1;                                                                  \
        e0 += kMmaChunk;                                                           \
        ca += cbytes;                                                              \
        cb += cbytes;                                                              \
        __syncwarp(); /* every lane is done with this stage */                     \
    }
        switch (type) {
            case T_Q4_K: MMA_UNIT(unit_k45<false>(sp, RS, 1, e0, sm, g, t, a0, a1)) break;
            case T_Q5_K: MMA_UNIT(unit_k45<true>(sp, RS, 1, e0, sm, g, t, a0, a1)) break;
            case T_Q6_K: MMA_UNIT(unit_q6k(sp, RS, 1, e0, ca & 15u, cb & 15u, sm, g, t, a0, a1)) break;
            default: MMA_UNIT(unit_q80(sp, RS, min(wsg.cb, wsg.nb_row - (e0 >> 5)), e0, ca & 15u, cb & 15u, sm, g, t, a0, a1)) break;
        }
#undef MMA_UNIT
        done += len;
        if (swiglu && cc.mat == 1) { au0 += r0; au1 += r1; } else { ag0 += r0; ag1 += r1; }

        // ---- tile finished (for this warp)? ----
        const int s = cc.s, tile = cc.tile;
        const bool row_end = (cc.chunk + len == p.chunks);
        const bool tile_done = (row_end && (!swiglu || cc.mat == 1)) || (done == n_units);
        if (row_end) cur_wrap(cc);
        if (!tile_done) continue;
        if (done == n_units) MMA_STAMP(5);

        // reduce the 4 lanes of a row group, then lane L holds logical row L (0..15 gate/plain, 16..31 up)
        ag0 += __shfl_xor_sync(0xffffffffu, ag0, 1); ag0 += __shfl_xor_sync(0xffffffffu, ag0, 2);
        ag1 += __shfl_xor_sync(0xffffffffu, ag1, 1); ag1 += __shfl_xor_sync(0xffffffffu, ag1, 2);
        if (swiglu) {
            au0 += __shfl_xor_sync(0xffffffffu, au0, 1); au0 += __shfl_xor_sync(0xffffffffu, au0, 2);
            au1 += __shfl_xor_sync(0xffffffffu, au1, 1); au1 += __shfl_xor_sync(0xffffffffu, au1, 2);
        }
        const int src = 4 * (lane & 7);
        const float vg0 = __shfl_sync(0xffffffffu, ag0, src), vg1 = __shfl_sync(0xffffffffu, ag1, src);
        const float vu0 = __shfl_sync(0xffffffffu, au0, src), vu1 = __shfl_sync(0xffffffffu, au1, src);
        const float v = (lane < 16) ? ((lane & 8) ? vg1 : vg0) : ((lane & 8) ? vu1 : vu0);
        ag0 = ag1 = au0 = au1 = 0.f;

        const int tu0 = p.seg[s].unit0 + tile * upt;
        if (u0 <= tu0 && tu0 + upt <= u1) {
            epilogue(s, tile, v);  // the whole tile is mine
        } else {
            const int slot = (u0 >= tu0) ? 0 : 1;  // tile is my first (slot 0) or starts inside my run (slot 1)
            s_part[warp][slot][lane] = v;
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
        float v = 0.f;
        for (int w = lo; w <= hi; w++) {  // fixed order: deterministic
            const int wu0 = cu0 + (w * cn) / nw, wu1 = cu0 + ((w + 1) * cn) / nw;
            if (wu1 == wu0) continue;     // a warp without units holds no piece
            v += s_part[w][(wu0 >= tu0) ? 0 : 1][lane];
        }
        if (tu0 < cu0 || tu0 + upt > cu1) {  // the tile straddles CTAs (tile_mode 0 only)
            const int tile_id = (s == 0 ? 0 : (s == 1 ? p.seg[0].n_tiles : p.seg[0].n_tiles + p.seg[1].n_tiles)) + tile;
            const int c_first = cta_of(tu0), c_last = cta_of(tu0 + upt - 1);
            p.part[((size_t)cta * 2 + ((cu0 >= tu0) ? 0 : 1)) * 32 + lane] = v;
            __syncwarp();
            unsigned int ticket = 0;
            if (lane == 0) ticket = atom_add_acq_rel(&p.tickets[tile_id], 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket != (unsigned)(c_last - c_first)) continue;  // not the last CTA
            v = 0.f;
            for (int c = c_first; c <= c_last; c++)
                v += ld_relaxed_gpu(&p.part[((size_t)c * 2 + ((cta_first(c) >= tu0) ? 0 : 1)) * 32 + lane]);
            if (lane == 0) p.tickets[tile_id] = 0;
        }
        epilogue(s, tile, v);
    }
    MMA_STAMP(6);
}

// Pull the first ring stages of launch/phase `p` towards L2 (fire-and-forget): called by the megakernel for the
// NEXT phase before it waits at a grid barrier, so the weight stream does not stop at the phase boundary.
template <int STAGES>
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
    const uint8_t* a = sg.w + (long long)min(tile * 16 + g, sg.n_rows - 1) * sg.row_bytes + off;
    const uint8_t* b = sg.w + (long long)min(tile * 16 + g + 8, sg.n_rows - 1) * sg.row_bytes + off;
    for (int o = 128 * t; o < bytes + 127; o += 512) {
        prefetch_l2(a + o);
        prefetch_l2(b + o);
    }
}

template <int STAGES>
__global__ void __launch_bounds__(kMmaMaxWarps * 32, 1) gemv_mma_kernel(const __grid_constant__ MParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_red[2 * kMmaMaxWarps];
    __shared__ float s_part[kMmaMaxWarps][2][32];   // pieces of tiles shared between warps of this CTA
    mma_gemv_cta<STAGES>(p, smem, s_red, s_part, [] {}, [] { pdl_launch_dependents(); pdl_wait(); }, false, true);
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
// must be set) and picks warps/stages so that x + the rings fit in shared memory.  Returns false if the
// launch is not eligible for this kernel (caller falls back to the CUDA-core kernel).
inline bool mma_plan(MParams& p, int n_sm, int want_warps, int want_stages, size_t smem_limit, MPlan& plan) {
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
        sg.n_tiles = (sg.n_rows + 15) / 16;
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
    p.stage_bytes = 16 * max_rs;
    const size_t xb = x_smem_bytes(p.K);
    int warps = std::max(4, std::min(want_warps, kMmaMaxWarps)), stages = std::max(2, std::min(want_stages, kMmaMaxStages));
    auto need = [&](int w, int st) { return xb + (size_t)w * st * p.stage_bytes + 16; };  // +16: funnel loads read one word past a piece
    while (need(warps, stages) > smem_limit) {
        if (stages > 2) stages--;
        else if (warps > 4) warps -= 2;
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
inline MmaKernel mma_kernel_for(int stages) {
    if (stages <= 2) return gemv_mma_kernel<2>;
    if (stages == 3) return gemv_mma_kernel<3>;
    return gemv_mma_kernel<4>;
}
inline cudaError_t mma_set_smem_limit(int bytes) {
    cudaError_t e = cudaFuncSetAttribute(gemv_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

}  // namespace b200
