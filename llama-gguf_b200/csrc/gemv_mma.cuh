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
//  * weights go HBM -> shared memory with cp.async.bulk (TMA bulk copies, SASS UBLKCP) into
//    PER-WARP mbarrier rings.  A unit is 16 rows x 512 (or 1024) elements, one bulk copy per
//    row; a warp produces and consumes its own ring, so there is no cross-warp hand-off and
//    the first stages are issued BEFORE griddepcontrol.wait (weights never depend on the
//    previous kernel of the token);
//  * dot products run on the tensor pipe (mma.sync.m16n8k16, SASS HMMA): quants become exact
//    fp16 integers with one LOP3/PRMT per two elements ((w & 0x000F000F) | 0x6400_6400 =
//    1024+q), x is split into fp16 hi + lo parts (x = hi + lo to 2^-22) that sit in two
//    columns of the B operand, accumulation is f32.  Block scales / mins are applied in f32
//    to per-sub-block sums (the reference's separated form, simd.rs:1006-1013); the integer
//    bias (1024, +32 for Q6_K, +128 for Q8_0) is removed with per-16-element sums of x.  The
//    8 columns of the MMA are used as 4 (hi, lo) pairs: the B operand is zero except in the
//    pair of the sub-block a lane's k-slots belong to, so each sub-block's sum lands in the
//    lane that decoded its scale and nothing is shuffled until a tile is finished;
//  * stream-K: all (tile, chunk) units of a launch are dealt evenly to the warps of a
//    persistent grid; tiles cut across warps are merged through a small scratch + ticket in
//    a fixed order (run-to-run deterministic).
//
// ~1.3 issue slots per weight (Q4_K) instead of ~6; see DESIGN.md for the budget.
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

constexpr int kMmaMaxWarps = 16;
constexpr int kMmaMaxStages = 4;

enum : int { ME_STORE = 0, ME_RESIDUAL = 1, ME_SWIGLU = 2, ME_SCALED_ACC = 3 };

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
    int row_stride;        // bytes between row slots of a ring stage
};

struct MParams {
    MSeg seg[3];
    int n_seg;
    int K;
    int chunk_elems;       // 512 or 1024
    int chunks;            // ceil(K / chunk_elems)
    int units_per_tile;    // chunks (2*chunks for ME_SWIGLU: gate chunks then up chunks)
    int total_units;
    int active_warps;      // min(grid*warps, total_units): every active warp owns >= 1 unit
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
    // stream-K merge scratch
    float* part;             // [grid*warps][2][32]
    unsigned int* tickets;   // [total logical tiles], zero between launches
    int* err;                // device error flag (watchdog)
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
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & mask) | orv
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (uint32_t)v;
}
__device__ __forceinline__ int lds_s8(uint32_t a) {
    int v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
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

constexpr uint32_t kMagic = 0x64006400u;  // half2(1024, 1024)

// ---------------------------------------------------------------- x in shared memory
// fp16 hi and lo parts (x ~= hi + lo), stored so that the 4 elements starting at e (e % 4 == 0)
// read as one 8-byte word give the m16n8k16 B fragment of a lane whose four k-slots come from one
// 32-bit word of quants: order [x0, x2, x1, x3] (bytes 0,2 -> k-slots 2t,2t+1; bytes 1,3 -> 2t+8,2t+9).
// xs16[i] = sum of float(hi)+float(lo) over elements 16i..16i+15.
struct XSmem {
    uint32_t xh, xl, xs;  // shared-space byte addresses
};
__host__ __device__ __forceinline__ int xperm(int e) { return (e & ~3) | (((e & 1) << 1) | ((e >> 1) & 1)); }
__host__ __device__ inline size_t x_smem_bytes(int K) { return ((size_t)4 * K + (size_t)(K >> 2) + 127) & ~(size_t)127; }

// All threads of the CTA.  K % 16 == 0.  Optional RMSNorm: y = (x * inv) * w (simd.rs:891-892).
__device__ __forceinline__ void stage_x_split(const float* __restrict__ x, const float* __restrict__ norm_w, float eps, int K,
                                              uint8_t* smem, float* red /*[kMmaMaxWarps]*/) {
    const int tid = threadIdx.x, nthr = blockDim.x, nwarp = nthr >> 5;
    float inv = 1.0f;
    if (norm_w) {
        float ss = 0.0f;
        for (int e = tid * 4; e < K; e += nthr * 4) {
            const float4 v = *reinterpret_cast<const float4*>(x + e);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        if ((tid & 31) == 0) red[tid >> 5] = ss;
        __syncthreads();
        float tot = 0.0f;
        for (int w = 0; w < nwarp; w++) tot += red[w];
        inv = 1.0f / sqrtf(tot / (float)K + eps);
    }
    __half* xh = reinterpret_cast<__half*>(smem);
    __half* xl = xh + K;
    float* xs = reinterpret_cast<float*>(smem + (size_t)4 * K);
    const int l16 = tid & 15;
    for (int e0 = (tid >> 4) * 16; e0 < K; e0 += (nthr >> 4) * 16) {
        const int e = e0 + l16;
        float v = x[e];
        if (norm_w) v = (v * inv) * norm_w[e];
        if (v > 65504.0f) v = 65504.0f;
        if (v < -65504.0f) v = -65504.0f;
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        xh[xperm(e)] = h;
        xl[xperm(e)] = l;
        float s = __half2float(h) + __half2float(l);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
        if (l16 == 0) xs[e0 >> 4] = s;
    }
}

// ---------------------------------------------------------------- per-type unit kernels
// sp = shared address of the stage (row slot r at sp + r*RS; the row's bytes start at +doff),
// nblk blocks, e0 = element index of the chunk's first element, g = lane>>2 (rows g and g+8),
// t = lane&3.  They add this unit's contribution for rows g and g+8 to acc0 / acc1 (partial over
// t: summed when the tile is finished).

__device__ __forceinline__ void k4_scales(const uint4& h, int t, float& dl, float& ml, float& dh, float& mh) {
    // get_scale_min_k4 (dequant.rs:213-225) for sub-blocks 2t (l) and 2t+1 (h)
    const float d = half_bits_to_float(h.x), dmin = half_bits_to_float(h.x >> 16);
    const int sh = 16 * (t & 1);
    const uint32_t A = (h.y >> sh) & 0xFFFFu, B = (h.z >> sh) & 0xFFFFu, C = (h.w >> sh) & 0xFFFFu;
    uint32_t scp, mnp;
    if (t < 2) {
        scp = A & 0x3F3Fu;
        mnp = B & 0x3F3Fu;
    } else {
        scp = (C & 0x0F0Fu) | ((A >> 2) & 0x3030u);
        mnp = ((C >> 4) & 0x0F0Fu) | ((B >> 2) & 0x3030u);
    }
    dl = d * (float)(scp & 0xFFu);
    dh = d * (float)(scp >> 8);
    ml = dmin * (float)(mnp & 0xFFu);
    mh = dmin * (float)(mnp >> 8);
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
    for (int b = 0; b < nblk; b++) {
        const uint32_t r0 = sp + g * RS + b * BB, r1 = r0 + 8 * RS;
        const uint4 h0 = lds128(r0), h1 = lds128(r1);
        uint4 qh0 = make_uint4(0u, 0u, 0u, 0u), qh1 = qh0;
        if (Q5) {
            qh0 = lds128(r0 + 16 + 16 * (t & 1));
            qh1 = lds128(r1 + 16 + 16 * (t & 1));
        }
        float cl[4] = {0.f, 0.f, 0.f, 0.f}, ch[4] = {0.f, 0.f, 0.f, 0.f};
        const int eb = e0 + b * 256;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint4 W0 = lds128(r0 + QS + 64 * c + 16 * t), W1 = lds128(r1 + QS + 64 * c + 16 * t);
            const int gp = 2 * c + (t >> 1);
            const bool act = lane_act && (c == c_act);
            const uint32_t xa = arr + 2u * (uint32_t)(eb + 64 * gp + 16 * (t & 1));
            uint4 bl[2], bh[2];
            bl[0] = bl[1] = bh[0] = bh[1] = make_uint4(0u, 0u, 0u, 0u);
            if (act) {
                bl[0] = lds128(xa);
                bl[1] = lds128(xa + 16);
                bh[0] = lds128(xa + 64);
                bh[1] = lds128(xa + 80);
            }
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
                const uint32_t wa8 = wa >> 8, wb8 = wb >> 8;
                mma16816(cl, lop3_and_or(wa, 0x000F000Fu, ml_a), lop3_and_or(wb, 0x000F000Fu, ml_b),
                         lop3_and_or(wa8, 0x000F000Fu, ml_a8), lop3_and_or(wb8, 0x000F000Fu, ml_b8), blx, bly);
                mma16816(ch, lop3_and_or(wa, 0x00F000F0u, mh_a), lop3_and_or(wb, 0x00F000F0u, mh_b),
                         lop3_and_or(wa8, 0x00F000F0u, mh_a8), lop3_and_or(wb8, 0x00F000F0u, mh_b8), bhx, bhy);
            }
        }
        // lane t owns sub-blocks 2t (low nibbles) and 2t+1 (high nibbles, carried x16) of this block
        float dl0, ml0, dh0, mh0, dl1, ml1, dh1, mh1;
        k4_scales(h0, t, dl0, ml0, dh0, mh0);
        k4_scales(h1, t, dl1, ml1, dh1, mh1);
        const uint32_t xsa = sm.xs + 4u * (uint32_t)((eb >> 4) + 4 * t);
        const float xsl = lds_f32(xsa) + lds_f32(xsa + 4), xsh = lds_f32(xsa + 8) + lds_f32(xsa + 12);
        acc0 += dl0 * (cl[0] + cl[1]) - (1024.0f * dl0 + ml0) * xsl + (dh0 * 0.0625f) * (ch[0] + ch[1]) - (64.0f * dh0 + mh0) * xsh;
        acc1 += dl1 * (cl[2] + cl[3]) - (1024.0f * dl1 + ml1) * xsl + (dh1 * 0.0625f) * (ch[2] + ch[3]) - (64.0f * dh1 + mh1) * xsh;
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
            const uint32_t xa = arr + 2u * (uint32_t)(eb + 128 * n + 16 * t);
            uint4 bl[2], bh[2];
            bl[0] = bl[1] = bh[0] = bh[1] = make_uint4(0u, 0u, 0u, 0u);
            if (act) {
                bl[0] = lds128(xa);
                bl[1] = lds128(xa + 16);
                bh[0] = lds128(xa + 128);
                bh[1] = lds128(xa + 144);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t lo0 = lop3_and_or(L0[i], 0x0F0F0F0Fu, (H0[i] << s_lo) & 0x30303030u);
                const uint32_t lo1 = lop3_and_or(L1[i], 0x0F0F0F0Fu, (H1[i] << s_lo) & 0x30303030u);
                const uint32_t hi0 = lop3_and_or(L0[i] >> 4, 0x0F0F0F0Fu, (H0[i] >> s_hi) & 0x30303030u);
                const uint32_t hi1 = lop3_and_or(L1[i] >> 4, 0x0F0F0F0Fu, (H1[i] >> s_hi) & 0x30303030u);
                const uint32_t blx = (i & 1) ? bl[i >> 1].z : bl[i >> 1].x, bly = (i & 1) ? bl[i >> 1].w : bl[i >> 1].y;
                const uint32_t bhx = (i & 1) ? bh[i >> 1].z : bh[i >> 1].x, bhy = (i & 1) ? bh[i >> 1].w : bh[i >> 1].y;
                mma16816(C[2 * n], __byte_perm(lo0, 0x64646464u, 0x4240), __byte_perm(lo1, 0x64646464u, 0x4240),
                         __byte_perm(lo0, 0x64646464u, 0x4341), __byte_perm(lo1, 0x64646464u, 0x4341), blx, bly);
                mma16816(C[2 * n + 1], __byte_perm(hi0, 0x64646464u, 0x4240), __byte_perm(hi1, 0x64646464u, 0x4240),
                         __byte_perm(hi0, 0x64646464u, 0x4341), __byte_perm(hi1, 0x64646464u, 0x4341), bhx, bhy);
            }
        }
        const float d0 = half_bits_to_float(lds16(r0 + 208)), d1 = half_bits_to_float(lds16(r1 + 208));
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int si = 4 * m + t;
            const float xs = lds_f32(sm.xs + 4u * (uint32_t)((eb >> 4) + si));
            const float s0 = (float)lds_s8(r0 + 192 + si), s1 = (float)lds_s8(r1 + 192 + si);
            acc0 += (d0 * s0) * ((C[m][0] + C[m][1]) - 1056.0f * xs);   // 1024 (fp16 magic) + 32 (Q6_K offset)
            acc1 += (d1 * s1) * ((C[m][2] + C[m][3]) - 1056.0f * xs);
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
                    uint2 bf = make_uint2(0u, 0u);
                    if (act) bf = lds64(arr + 2u * (uint32_t)(e0 + 32 * b + 16 * m + 4 * t));
                    mma16816(C, __byte_perm(w0, 0x64646464u, 0x4240), __byte_perm(w1, 0x64646464u, 0x4240),
                             __byte_perm(w0, 0x64646464u, 0x4341), __byte_perm(w1, 0x64646464u, 0x4341), bf.x, bf.y);
                }
            }
        }
        const int b = b4 + t;  // lane t owns block b4 + t
        if (b < nblk) {
            const uint32_t xa = sm.xs + 4u * (uint32_t)(((e0 + 32 * b) >> 4));
            const float xs = lds_f32(xa) + lds_f32(xa + 4);
            acc0 += half_bits_to_float(lds16(row0 + b * 34)) * ((C[0] + C[1]) - 1152.0f * xs);  // 1024 + 128
            acc1 += half_bits_to_float(lds16(row1 + b * 34)) * ((C[2] + C[3]) - 1152.0f * xs);
        }
    }
}

// ---------------------------------------------------------------- the kernel
__device__ __forceinline__ float mma_silu(float x) { return x / (1.0f + expf(-x)); }

__host__ __device__ inline int mma_chunk_blocks(int type, int chunk_elems) { return chunk_elems / type_block_elems(type); }

// warp that owns unit u when U units are dealt to W warps as [floor(i*U/W), floor((i+1)*U/W))  (W <= U)
__device__ __forceinline__ int mma_owner(long long u, long long U, long long W) { return (int)(((u + 1) * W - 1) / U); }

__global__ void __launch_bounds__(kMmaMaxWarps * 32, 1) gemv_mma_kernel(const MParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_red[kMmaMaxWarps];
    __shared__ __align__(8) unsigned long long s_bars[kMmaMaxWarps * kMmaMaxStages];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int K = p.K;
    const uint32_t sbase = smem_u32(smem);
    XSmem sm;
    sm.xh = sbase;
    sm.xl = sbase + 2u * K;
    sm.xs = sbase + 4u * K;
    const uint32_t ring = sbase + (uint32_t)x_smem_bytes(K) + (uint32_t)warp * p.stages * p.stage_bytes;
    const uint32_t wbar = smem_u32(&s_bars[warp * kMmaMaxStages]);

    if (lane == 0)
        for (int s = 0; s < p.stages; s++) mbar_init(wbar + 8 * s, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    // interleaved warp numbering: consecutive global warps sit on different SMs
    const long long U = p.total_units, W = p.active_warps;
    const long long gw = (long long)warp * gridDim.x + blockIdx.x;
    const bool active = gw < W;
    const int u0 = active ? (int)(gw * U / W) : 0, u1 = active ? (int)((gw + 1) * U / W) : 0;
    const int n_units = u1 - u0;
    long long eoff = 0;  // MoE expert index (valid after pdl_wait)

    // unit -> (segment, logical tile, chunk)
    auto decode = [&](int u, int& s, int& tile, int& chunk) {
        s = (p.n_seg > 2 && u >= p.seg[2].unit0) ? 2 : (p.n_seg > 1 && p.epi != ME_SWIGLU && u >= p.seg[1].unit0) ? 1 : 0;
        const int local = u - p.seg[s].unit0;
        tile = local / p.units_per_tile;
        chunk = local - tile * p.units_per_tile;
    };
    // byte offset (within its row slot) at which the bytes of row `row` of segment sg, chunk `chunk` start
    auto row_src = [&](const MSeg& sg, int tile, int r, int chunk) -> const uint8_t* {
        const int row = min(tile * 16 + r, sg.n_rows - 1);
        const int cb = mma_chunk_blocks(sg.type, p.chunk_elems), bb = type_block_bytes(sg.type);
        return sg.w + eoff * sg.expert_stride + (long long)row * sg.row_bytes + (long long)chunk * cb * bb;
    };
    // issue the bulk copies of unit u into ring stage st (lanes 0..15: one row each)
    auto issue = [&](int u, int st) {
        int s, tile, chunk;
        decode(u, s, tile, chunk);
        if (p.epi == ME_SWIGLU && chunk >= p.chunks) { s = 1; chunk -= p.chunks; }
        const MSeg& sg = p.seg[s];
        if (lane < 16) {
            const int cb = mma_chunk_blocks(sg.type, p.chunk_elems), bb = type_block_bytes(sg.type), be = type_block_elems(sg.type);
            const int nblk = min(cb, K / be - chunk * cb);
            const uint8_t* src = row_src(sg, tile, lane, chunk);
            const uint32_t doff = (uint32_t)((uintptr_t)src & 15u);
            const uint32_t bytes = (doff + (uint32_t)(nblk * bb) + 15u) & ~15u;
            const uint32_t bar = wbar + 8 * st;
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(ring + (uint32_t)st * p.stage_bytes + (uint32_t)lane * sg.row_stride, src - doff, bytes, bar);
        }
    };

    const int pre = min(p.stages - 1, n_units);
    if (!p.expert_sel)  // dense weights never depend on a predecessor: start streaming before the PDL wait
        for (int k = 0; k < pre; k++) issue(u0 + k, k);

    pdl_launch_dependents();
    pdl_wait();

    if (p.expert_sel) {
        eoff = (long long)p.expert_sel[p.expert_slot];
        for (int k = 0; k < pre; k++) issue(u0 + k, k);
    }
    stage_x_split(p.x, p.norm_w, p.eps, K, smem, s_red);
    __syncthreads();

    float ag0 = 0.f, ag1 = 0.f, au0 = 0.f, au1 = 0.f;
    bool ok = true;
    for (int k = 0; k < n_units && ok; k++) {
        const int u = u0 + k;
        if (k + p.stages - 1 < n_units) {
            __syncwarp();
            issue(u + p.stages - 1, (k + p.stages - 1) % p.stages);
        }
        const int st = k % p.stages;
        ok = mbar_wait(wbar + 8 * st, (uint32_t)((k / p.stages) & 1), p.err);
        if (!ok) break;

        int s, tile, chunk;
        decode(u, s, tile, chunk);
        bool is_up = false;
        int ws = s;
        if (p.epi == ME_SWIGLU && chunk >= p.chunks) { is_up = true; ws = 1; chunk -= p.chunks; }
        const MSeg& wsg = p.seg[ws];
        const int type = wsg.type;
        const int cb = mma_chunk_blocks(type, p.chunk_elems), be = type_block_elems(type);
        const int nblk = min(cb, K / be - chunk * cb);
        const int e0 = chunk * p.chunk_elems;
        const uint32_t sp = ring + (uint32_t)st * p.stage_bytes;
        const uint32_t RS = (uint32_t)wsg.row_stride;
        float a0 = 0.f, a1 = 0.f;
        switch (type) {
            case T_Q4_K: unit_k45<false>(sp, RS, nblk, e0, sm, g, t, a0, a1); break;
            case T_Q5_K: unit_k45<true>(sp, RS, nblk, e0, sm, g, t, a0, a1); break;
            case T_Q6_K: {
                const uint32_t d0 = (uint32_t)((uintptr_t)row_src(wsg, tile, g, chunk) & 15u);
                const uint32_t d1 = (uint32_t)((uintptr_t)row_src(wsg, tile, g + 8, chunk) & 15u);
                unit_q6k(sp, RS, nblk, e0, d0, d1, sm, g, t, a0, a1);
                break;
            }
            default: {
                const uint32_t d0 = (uint32_t)((uintptr_t)row_src(wsg, tile, g, chunk) & 15u);
                const uint32_t d1 = (uint32_t)((uintptr_t)row_src(wsg, tile, g + 8, chunk) & 15u);
                unit_q80(sp, RS, nblk, e0, d0, d1, sm, g, t, a0, a1);
                break;
            }
        }
        if (is_up) { au0 += a0; au1 += a1; } else { ag0 += a0; ag1 += a1; }

        // ---- tile finished (for this warp)? ----
        int s2 = -1, tile2 = -1, chunk2;
        if (k + 1 < n_units) decode(u + 1, s2, tile2, chunk2);
        if (s2 == s && tile2 == tile) continue;

        // reduce the 4 lanes of a row group, then lane L holds logical row L (0..15 gate/plain, 16..31 up)
        ag0 += __shfl_xor_sync(0xffffffffu, ag0, 1); ag0 += __shfl_xor_sync(0xffffffffu, ag0, 2);
        ag1 += __shfl_xor_sync(0xffffffffu, ag1, 1); ag1 += __shfl_xor_sync(0xffffffffu, ag1, 2);
        au0 += __shfl_xor_sync(0xffffffffu, au0, 1); au0 += __shfl_xor_sync(0xffffffffu, au0, 2);
        au1 += __shfl_xor_sync(0xffffffffu, au1, 1); au1 += __shfl_xor_sync(0xffffffffu, au1, 2);
        const int src = 4 * (lane & 7);
        const float vg0 = __shfl_sync(0xffffffffu, ag0, src), vg1 = __shfl_sync(0xffffffffu, ag1, src);
        const float vu0 = __shfl_sync(0xffffffffu, au0, src), vu1 = __shfl_sync(0xffffffffu, au1, src);
        float v = (lane < 16) ? ((lane & 8) ? vg1 : vg0) : ((lane & 8) ? vu1 : vu0);
        ag0 = ag1 = au0 = au1 = 0.f;

        // stream-K merge: which warps hold pieces of this tile?
        const long long tu0 = (long long)p.seg[s].unit0 + (long long)tile * p.units_per_tile;
        const int w_first = mma_owner(tu0, U, W), w_last = mma_owner(tu0 + p.units_per_tile - 1, U, W);
        const int tile_id = (s == 0 ? 0 : (s == 1 ? p.seg[0].n_tiles : p.seg[0].n_tiles + p.seg[1].n_tiles)) + tile;
        if (w_last != w_first) {
            const int slot = ((long long)u0 >= tu0) ? 0 : 1;  // tile is my first (slot 0) or my last (slot 1)
            p.part[((size_t)gw * 2 + slot) * 32 + lane] = v;
            __threadfence();
            __syncwarp();
            unsigned int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&p.tickets[tile_id], 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket != (unsigned)(w_last - w_first)) continue;  // not the last piece
            __threadfence();
            v = 0.f;
            for (int wi = w_first; wi <= w_last; wi++) {
                const long long wu0 = (long long)wi * U / W;
                const int sl = (wu0 >= tu0) ? 0 : 1;
                v += __ldcg(&p.part[((size_t)wi * 2 + sl) * 32 + lane]);
            }
            if (lane == 0) p.tickets[tile_id] = 0;
        }

        // ---- epilogue: lane L < 16 owns row j of segment s ----
        const MSeg& sg = p.seg[s];
        const int j = tile * 16 + (lane & 15);
        const bool valid = (lane < 16) && (j < sg.n_rows);
        float val = v;
        if (p.epi == ME_SWIGLU) {
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
            sg.out[j] = val;
        }
    }
}

// ---------------------------------------------------------------- host-side launch planning
inline bool mma_type_ok(int type) { return type == T_Q4_K || type == T_Q5_K || type == T_Q6_K || type == T_Q8_0; }

// bytes between row slots: room for the chunk (+15 bytes of source misalignment, +4 of funnel over-read),
// residue mod 128 chosen for conflict-free fragment loads (LDS.128: 64; 32-bit loads: odd multiple of 16)
inline int mma_row_stride(int type, int chunk_elems) {
    const int cb = mma_chunk_blocks(type, chunk_elems), bb = type_block_bytes(type);
    const bool aligned = (type == T_Q4_K || type == T_Q5_K);
    int rs = aligned ? cb * bb : ((cb * bb + 15 + 4 + 15) & ~15);
    for (;; rs += 16) {
        const int m = rs & 127;
        if (aligned ? (m == 64) : ((m & 15) == 0 && ((m >> 4) & 1))) return rs;
    }
}

struct MPlan {
    int grid, warps, stages;
    size_t smem;
};

// Fills the derived fields of p (segments' w/out/bias/row_bytes/expert_stride/type/n_rows, n_seg, K, epi
// must be set) and picks warps/stages so that x + the rings fit in shared memory.  Returns false if the
// launch is not eligible for this kernel (caller falls back to the CUDA-core kernel).
inline bool mma_plan(MParams& p, int n_sm, int chunk_elems, int want_warps, int want_stages, size_t smem_limit, MPlan& plan) {
    if (p.K % 32) return false;
    if (p.epi == ME_SWIGLU && (p.n_seg != 2 || p.seg[0].n_rows != p.seg[1].n_rows)) return false;
    int max_rs = 0;
    for (int s = 0; s < p.n_seg; s++) {
        MSeg& sg = p.seg[s];
        if (!mma_type_ok(sg.type)) return false;
        if (p.K % type_block_elems(sg.type)) return false;
        if ((sg.type == T_Q4_K || sg.type == T_Q5_K) && ((sg.row_bytes & 15) || (sg.expert_stride & 15))) return false;
        if ((sg.row_bytes & 1) || (sg.expert_stride & 1)) return false;
        if (chunk_elems % type_block_elems(sg.type)) return false;
        sg.row_stride = mma_row_stride(sg.type, chunk_elems);
        sg.n_tiles = (sg.n_rows + 15) / 16;
        max_rs = std::max(max_rs, sg.row_stride);
    }
    p.chunk_elems = chunk_elems;
    p.chunks = (p.K + chunk_elems - 1) / chunk_elems;
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
    int warps = std::min(want_warps, kMmaMaxWarps), stages = std::min(want_stages, kMmaMaxStages);
    auto need = [&](int w, int st) { return xb + (size_t)w * st * p.stage_bytes; };
    while (need(warps, stages) > smem_limit) {
        if (stages > 2) stages--;
        else if (warps > 4) warps -= 2;
        else return false;
    }
    p.stages = stages;
    plan.warps = warps;
    plan.stages = stages;
    plan.smem = need(warps, stages);
    const long long slots = (long long)n_sm * warps;
    plan.grid = (int)std::min<long long>(n_sm, (p.total_units + warps - 1) / warps);
    if (plan.grid < 1) plan.grid = 1;
    p.active_warps = (int)std::min<long long>((long long)plan.grid * warps, p.total_units);
    (void)slots;
    return true;
}

}  // namespace b200
