// gemv2.cuh — the production dequant-GEMV for batch-1 decode (Q4_K / Q5_K / Q6_K / Q8_0).
//
// Same contract as gemv.cuh (y = deq(W) x on the untouched GGUF super-block layout, fused
// norm / bias / residual / SwiGLU / RoPE+KV-write / MoE-accumulate), re-designed around what
// ncu showed for the first version (profiles/r01_v1_*): ~6 issue slots per weight element
// made the CUDA-core kernel instruction-bound at 1.2 TB/s.  Here:
//
//  * weights move HBM -> shared memory with cp.async.bulk (TMA bulk copies, SASS UBLKCP)
//    into PER-WARP mbarrier rings: a "unit" is 16 rows x 512 elements (2 K-quant blocks per
//    row, one bulk copy per row), so DRAM sees long contiguous bursts and no LSU address
//    work; the first ring stages are issued BEFORE griddepcontrol.wait, i.e. while the
//    previous kernel of the token is still draining (weights never depend on it);
//  * the dot products run on the tensor pipe: nibbles become exact fp16 integers with one
//    LOP3 per two elements ((w & 0x000F000F) | 0x6400_6400 = 1024+q), x is split into
//    fp16 hi + lo parts (x = hi + lo to ~2^-22) that sit in two columns of the B operand of
//    mma.sync.m16n8k16, accumulation is f32.  Block scales / mins are applied in f32 to the
//    per-sub-block sums (the reference's separated form, simd.rs:1006-1013), the integer
//    bias is removed with per-16-element sums of x.  Column pairs route each sub-block's sum
//    to the lane that decoded its scale, so no shuffles are needed until a tile is finished;
//  * work is split stream-K style: all (tile, chunk) units of a launch are dealt evenly to
//    every warp of a persistent grid (2 CTAs/SM); tiles cut across warps are merged through a
//    small scratch + ticket, in a fixed order (deterministic);
//  * the epilogue of the producer writes the NEXT GEMV's input already prepared (fp16 hi/lo
//    in fragment order, 16-element sums, sum-of-squares partials for the RMSNorm), so no
//    kernel re-reads and re-converts x: RMSNorm becomes a scalar applied in the epilogue.
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

// ---------------------------------------------------------------- prepared activations
// element e of the vector lives at half-index (e & ~3) | kPerm[e & 3]: [x0, x2, x1, x3]
// is the order the m16n8k16 B fragment wants when a lane's four k-slots come from one
// 32-bit word of quants (bytes 0,2 -> slots 2t,2t+1; bytes 1,3 -> slots 2t+8,2t+9).
struct XPrep {
    __half* hi;    // [K]
    __half* lo;    // [K]
    float* xs16;   // [K/16] sums of (float(hi)+float(lo)) over 16 consecutive elements
    float* ssq;    // [K/16] sums of squares of the f32 values (RMSNorm of the consumer); may be null
};
__host__ __device__ __forceinline__ int xperm(int e) { return (e & ~3) | (((e & 1) << 1) | ((e >> 1) & 1)); }

// Standalone preparation (embedding output, per-op API, fallbacks): one CTA.
__global__ void __launch_bounds__(256) xprep_kernel(const float* __restrict__ x, const float* __restrict__ w, int K, XPrep y) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 15;
    for (int e0 = (threadIdx.x >> 4) * 16; e0 < K; e0 += (blockDim.x >> 4) * 16) {
        const int e = e0 + lane;
        float v = x[e];
        float yv = w ? v * w[e] : v;
        yv = fminf(fmaxf(yv, -65504.0f), 65504.0f);
        const __half h = __float2half_rn(yv);
        const __half l = __float2half_rn(yv - __half2float(h));
        y.hi[xperm(e)] = h;
        y.lo[xperm(e)] = l;
        float s = __half2float(h) + __half2float(l);
        float q = v * v;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o, 16);
            q += __shfl_xor_sync(0xffffffffu, q, o, 16);
        }
        if (lane == 0) {
            y.xs16[e0 >> 4] = s;
            if (y.ssq) y.ssq[e0 >> 4] = q;
        }
    }
}

// ---------------------------------------------------------------- parameters
constexpr int kG2Warps = 8;
constexpr int kG2Threads = kG2Warps * 32;
constexpr int kG2ChunkElems = 512;
constexpr int kG2MaxStages = 4;

enum : int { E2_STORE = 0, E2_RESIDUAL = 1, E2_SWIGLU = 2, E2_QKV_ROPE = 3, E2_SCALED_ACC = 4 };

struct G2Seg {
    const uint8_t* w;
    float* out;            // optional f32 output [n_rows]
    const float* bias;     // optional
    long long row_bytes;
    long long expert_stride;
    int type;
    int n_rows;
    int n_tiles;           // ceil(n_rows / 16)
    int unit0;             // first unit of this segment in the launch
};

struct G2Params {
    G2Seg seg[3];
    int n_seg;
    int K;
    int chunks;            // ceil(K / 512)
    int units_per_tile;    // chunks (2*chunks for E2_SWIGLU: gate chunks then up chunks)
    int total_units;
    int stages;
    int stage_bytes;       // 16 * row stride, max over segments
    XPrep x;
    int use_norm;          // scale the result by rsqrt(mean(x^2) + eps) (RMSNorm folded out of the GEMV)
    float eps;
    int epi;
    const float* residual;
    XPrep y;               // y.hi == nullptr: no prepared output
    const float* y_w;      // optional element-wise weight folded into y (the consumer's RMSNorm weight)
    // E2_QKV_ROPE (normal-style pairs only): seg0 = q (rotated, stored to seg.out), seg1 = k, seg2 = v -> caches
    const float* freq;
    const int* pos;
    float* k_cache;
    float* v_cache;
    int hd, max_seq;
    float rope_scale;
    // MoE
    const int* expert_sel;
    const float* expert_wt;
    int expert_slot;
    // stream-K merge scratch
    float* part;             // [total warps][2][32]
    unsigned int* tickets;   // [total logical tiles], zero between launches
    int* err;                // device error flag (watchdog)
};

__host__ __device__ inline int g2_chunk_blocks(int type) { return type == T_Q8_0 ? 16 : 2; }
__host__ __device__ inline int g2_row_stride(int type) {
    switch (type) {
        case T_Q4_K: return 288;
        case T_Q5_K: return 352;
        case T_Q6_K: return 464;
        case T_Q8_0: return 560;
    }
    return 0;
}
__host__ inline bool g2_eligible(int type, long long k, long long row_bytes) {
    if (!(type == T_Q4_K || type == T_Q5_K || type == T_Q6_K || type == T_Q8_0)) return false;
    return k % 256 == 0 && row_bytes % 16 == 0;
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
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
    for (int it = 0; it < (1 << 22); it++)
        if (mbar_try_wait(bar, parity)) return true;
    if (err) atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
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
// 32 bits at a 2-byte-aligned shared address
__device__ __forceinline__ uint32_t lds32_a2(uint32_t a) { return lds16(a) | (lds16(a + 2) << 16); }

constexpr uint32_t kMagic = 0x64006400u;  // half2(1024, 1024)

struct G2Smem {
    uint32_t xh, xl, xs;  // shared-space byte addresses of x_hi, x_lo (halves), xs16 (floats)
};

// B fragment (two k-slot pairs of the column this lane owns) for the 4 elements starting at `e` (e % 4 == 0)
__device__ __forceinline__ uint2 load_bfrag(uint32_t arr, int e, bool act) {
    uint2 b = make_uint2(0u, 0u);
    if (act) b = lds64(arr + 2u * (uint32_t)e);
    return b;
}

// ---------------------------------------------------------------- per-type unit kernels
// All take: sp = shared address of the stage (row r at sp + r*RS), nblk blocks, e0 = element index of the
// chunk's first element, lane coordinates g = lane>>2 (row g and g+8), t = lane&3.  They add this unit's
// contribution for rows g and g+8 to acc0 / acc1 (partial over t: summed when the tile is finished).

__device__ __forceinline__ void k4_scales(const uint4& h, int t, float& dl, float& ml, float& dh, float& mh) {
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

template <bool Q5>
__device__ __forceinline__ void unit_k45(uint32_t sp, int nblk, int e0, const G2Smem& sm, int g, int t, float& acc0,
                                         float& acc1) {
    constexpr int BB = Q5 ? 176 : 144, QS = Q5 ? 48 : 16, RS = Q5 ? 352 : 288;
    const uint32_t arr = (g & 1) ? sm.xl : sm.xh;
    for (int b = 0; b < nblk; b++) {
        const uint32_t r0 = sp + g * RS + b * BB, r1 = r0 + 8 * RS;
        const uint4 h0 = lds128(r0), h1 = lds128(r1);
        float dl0, ml0, dh0, mh0, dl1, ml1, dh1, mh1;
        k4_scales(h0, t, dl0, ml0, dh0, mh0);
        k4_scales(h1, t, dl1, ml1, dh1, mh1);
        float cl[4] = {0.f, 0.f, 0.f, 0.f}, ch[4] = {0.f, 0.f, 0.f, 0.f};
        const int eb = e0 + b * 256;
#pragma unroll
        for (int gp = 0; gp < 4; gp++) {
            const uint2 w0 = lds64(r0 + QS + 32 * gp + 8 * t), w1 = lds64(r1 + QS + 32 * gp + 8 * t);
            uint2 q0 = make_uint2(0u, 0u), q1 = make_uint2(0u, 0u);
            if (Q5) {
                q0 = lds64(r0 + 16 + 8 * t);  // qh bytes of the same 8 byte positions (shared by all gp)
                q1 = lds64(r1 + 16 + 8 * t);
            }
            const bool act = (g >> 1) == gp;
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const uint32_t wa = m ? w0.y : w0.x, wb = m ? w1.y : w1.x;
                const int e = eb + 64 * gp + 8 * t + 4 * m;
                const uint2 bl = load_bfrag(arr, e, act), bh = load_bfrag(arr, e + 32, act);
                uint32_t ml_a = kMagic, ml_b = kMagic, ml_a8 = kMagic, ml_b8 = kMagic;  // low-group or-values
                uint32_t mh_a = kMagic, mh_b = kMagic, mh_a8 = kMagic, mh_b8 = kMagic;  // high-group or-values
                if (Q5) {  // 5th bit: +16 for the low group (bit 4), +256 (= 16*16) for the x16-scaled high group (bit 8)
                    const uint32_t ha = m ? q0.y : q0.x, hb = m ? q1.y : q1.x;
                    const uint32_t la = ha >> (2 * gp), lb = hb >> (2 * gp);  // bit0 of each byte = low-group bit
                    ml_a = lop3_and_or(la << 4, 0x00100010u, kMagic);
                    ml_b = lop3_and_or(lb << 4, 0x00100010u, kMagic);
                    ml_a8 = lop3_and_or(la >> 4, 0x00100010u, kMagic);
                    ml_b8 = lop3_and_or(lb >> 4, 0x00100010u, kMagic);
                    mh_a = lop3_and_or(la << 7, 0x01000100u, kMagic);  // bit1 of each byte = high-group bit -> bit 8
                    mh_b = lop3_and_or(lb << 7, 0x01000100u, kMagic);
                    mh_a8 = lop3_and_or(la >> 1, 0x01000100u, kMagic);
                    mh_b8 = lop3_and_or(lb >> 1, 0x01000100u, kMagic);
                }
                const uint32_t wa8 = wa >> 8, wb8 = wb >> 8;
                mma16816(cl, lop3_and_or(wa, 0x000F000Fu, ml_a), lop3_and_or(wb, 0x000F000Fu, ml_b),
                         lop3_and_or(wa8, 0x000F000Fu, ml_a8), lop3_and_or(wb8, 0x000F000Fu, ml_b8), bl.x, bl.y);
                mma16816(ch, lop3_and_or(wa, 0x00F000F0u, mh_a), lop3_and_or(wb, 0x00F000F0u, mh_b),
                         lop3_and_or(wa8, 0x00F000F0u, mh_a8), lop3_and_or(wb8, 0x00F000F0u, mh_b8), bh.x, bh.y);
            }
        }
        // lane t owns sub-blocks 2t (low nibbles) and 2t+1 (high nibbles, carried x16) of this block
        const uint32_t xa = sm.xs + 4u * (uint32_t)((eb >> 4) + 4 * t);
        const float xsl = lds_f32(xa) + lds_f32(xa + 4), xsh = lds_f32(xa + 8) + lds_f32(xa + 12);
        const float dh0s = dh0 * 0.0625f, dh1s = dh1 * 0.0625f;
        acc0 += dl0 * (cl[0] + cl[1]) - (1024.0f * dl0 + ml0) * xsl + dh0s * (ch[0] + ch[1]) - (64.0f * dh0 + mh0) * xsh;
        acc1 += dl1 * (cl[2] + cl[3]) - (1024.0f * dl1 + ml1) * xsl + dh1s * (ch[2] + ch[3]) - (64.0f * dh1 + mh1) * xsh;
    }
}

// Q6_K: 16 scale groups of 16 elements per block.  Group (n, c, h) -> accumulator set c, column pair 2n+h.
__device__ __forceinline__ void unit_q6k(uint32_t sp, int nblk, int e0, int doff, const G2Smem& sm, int g, int t,
                                         float& acc0, float& acc1) {
    constexpr int RS = 464;
    const uint32_t arr = (g & 1) ? sm.xl : sm.xh;
    for (int b = 0; b < nblk; b++) {
        const uint32_t r0 = sp + g * RS + doff + b * 210, r1 = r0 + 8 * RS;
        float C[4][4];
#pragma unroll
        for (int c = 0; c < 4; c++) C[c][0] = C[c][1] = C[c][2] = C[c][3] = 0.f;
        const int eb = e0 + b * 256;
#pragma unroll
        for (int nh = 0; nh < 4; nh++) {
            const int n = nh >> 1, h = nh & 1;
            const uint32_t o = 64 * n + 16 * h + 4 * t;
            const uint32_t A0 = lds32_a2(r0 + o), B0 = lds32_a2(r0 + o + 32), H0 = lds32_a2(r0 + 128 + 32 * n + 16 * h + 4 * t);
            const uint32_t A1 = lds32_a2(r1 + o), B1 = lds32_a2(r1 + o + 32), H1 = lds32_a2(r1 + 128 + 32 * n + 16 * h + 4 * t);
            uint32_t q0[4], q1[4];
            q0[0] = lop3_and_or(A0, 0x0F0F0F0Fu, (H0 << 4) & 0x30303030u);
            q0[1] = lop3_and_or(B0, 0x0F0F0F0Fu, (H0 << 2) & 0x30303030u);
            q0[2] = lop3_and_or(A0 >> 4, 0x0F0F0F0Fu, H0 & 0x30303030u);
            q0[3] = lop3_and_or(B0 >> 4, 0x0F0F0F0Fu, (H0 >> 2) & 0x30303030u);
            q1[0] = lop3_and_or(A1, 0x0F0F0F0Fu, (H1 << 4) & 0x30303030u);
            q1[1] = lop3_and_or(B1, 0x0F0F0F0Fu, (H1 << 2) & 0x30303030u);
            q1[2] = lop3_and_or(A1 >> 4, 0x0F0F0F0Fu, H1 & 0x30303030u);
            q1[3] = lop3_and_or(B1 >> 4, 0x0F0F0F0Fu, (H1 >> 2) & 0x30303030u);
            const bool act = (g >> 1) == nh;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint2 bf = load_bfrag(arr, eb + 128 * n + 32 * c + 16 * h + 4 * t, act);
                mma16816(C[c], __byte_perm(q0[c], 0x64646464u, 0x4240), __byte_perm(q1[c], 0x64646464u, 0x4240),
                         __byte_perm(q0[c], 0x64646464u, 0x4341), __byte_perm(q1[c], 0x64646464u, 0x4341), bf.x, bf.y);
            }
        }
        // lane t owns column pair t = (n = t>>1, h = t&1): groups (n, c, h), scale index 8n + 2c + h
        const int n = t >> 1, h = t & 1;
        const float d0 = half_bits_to_float(lds16(r0 + 208)), d1 = half_bits_to_float(lds16(r1 + 208));
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int si = 8 * n + 2 * c + h;
            const float xs = lds_f32(sm.xs + 4u * (uint32_t)((eb >> 4) + si));
            const float s0 = (float)lds_s8(r0 + 192 + si), s1 = (float)lds_s8(r1 + 192 + si);
            acc0 += (d0 * s0) * ((C[c][0] + C[c][1]) - 1056.0f * xs);   // 1024 (fp16 magic) + 32 (Q6_K offset)
            acc1 += (d1 * s1) * ((C[c][2] + C[c][3]) - 1056.0f * xs);
        }
    }
}

// Q8_0: 32-element blocks; blocks 4i..4i+3 share one accumulator set through the column pairs.
__device__ __forceinline__ void unit_q80(uint32_t sp, int nblk, int e0, const G2Smem& sm, int g, int t, float& acc0,
                                         float& acc1) {
    constexpr int RS = 560;
    const uint32_t arr = (g & 1) ? sm.xl : sm.xh;
    for (int b4 = 0; b4 < nblk; b4 += 4) {
        float C[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int bi = 0; bi < 4; bi++) {
            const int b = b4 + bi;
            if (b < nblk) {  // warp-uniform
                const uint32_t r0 = sp + g * RS + b * 34 + 2, r1 = r0 + 8 * RS;
                const bool act = (g >> 1) == bi;
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    const uint32_t w0 = lds32_a2(r0 + 16 * m + 4 * t) ^ 0x80808080u;  // int8 -> biased uint8
                    const uint32_t w1 = lds32_a2(r1 + 16 * m + 4 * t) ^ 0x80808080u;
                    const uint2 bf = load_bfrag(arr, e0 + 32 * b + 16 * m + 4 * t, act);
                    mma16816(C, __byte_perm(w0, 0x64646464u, 0x4240), __byte_perm(w1, 0x64646464u, 0x4240),
                             __byte_perm(w0, 0x64646464u, 0x4341), __byte_perm(w1, 0x64646464u, 0x4341), bf.x, bf.y);
                }
            }
        }
        const int b = b4 + t;  // lane t owns block b4 + t
        if (b < nblk) {
            const uint32_t r0 = sp + g * RS + b * 34, r1 = r0 + 8 * RS;
            const uint32_t xa = sm.xs + 4u * (uint32_t)(((e0 + 32 * b) >> 4));
            const float xs = lds_f32(xa) + lds_f32(xa + 4);
            acc0 += half_bits_to_float(lds16(r0)) * ((C[0] + C[1]) - 1152.0f * xs);  // 1024 + 128
            acc1 += half_bits_to_float(lds16(r1)) * ((C[2] + C[3]) - 1152.0f * xs);
        }
    }
}

// ---------------------------------------------------------------- the kernel
__device__ __forceinline__ float g2_silu(float x) { return x / (1.0f + expf(-x)); }

// warp that owns unit u when U units are dealt to W warps as [floor(i*U/W), floor((i+1)*U/W))
__device__ __forceinline__ int g2_owner(long long u, long long U, long long W) { return (int)(((u + 1) * W - 1) / U); }

__global__ void __launch_bounds__(kG2Threads, 2) gemv2_kernel(const G2Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_inv;
    __shared__ __align__(8) unsigned long long s_bars[1 + kG2Warps * kG2MaxStages];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int K = p.K;
    const uint32_t sbase = smem_u32(smem);
    G2Smem sm;
    sm.xh = sbase;
    sm.xl = sbase + 2u * K;
    sm.xs = sbase + 4u * K;
    const uint32_t xbytes = 4u * K + (uint32_t)(K >> 2);
    const uint32_t ring = sbase + ((xbytes + 127u) & ~127u) + (uint32_t)warp * p.stages * p.stage_bytes;
    const uint32_t xbar = smem_u32(&s_bars[0]);
    const uint32_t wbar = smem_u32(&s_bars[1 + warp * kG2MaxStages]);

    if (threadIdx.x == 0) mbar_init(xbar, 1);
    if (lane == 0)
        for (int s = 0; s < p.stages; s++) mbar_init(wbar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const long long U = p.total_units, W = (long long)gridDim.x * kG2Warps;
    const long long gw = (long long)blockIdx.x * kG2Warps + warp;
    const int u0 = (int)(gw * U / W), u1 = (int)((gw + 1) * U / W);
    const int n_units = u1 - u0;
    long long eoff = 0;  // MoE expert row offset (valid after pdl_wait)

    // unit -> (segment, logical tile, chunk)
    auto decode = [&](int u, int& s, int& tile, int& chunk) {
        s = (p.n_seg > 2 && u >= p.seg[2].unit0) ? 2 : (p.n_seg > 1 && p.epi != E2_SWIGLU && u >= p.seg[1].unit0) ? 1 : 0;
        const int local = u - p.seg[s].unit0;
        tile = local / p.units_per_tile;
        chunk = local - tile * p.units_per_tile;
    };
    // issue the bulk copies of unit u into ring stage st (lanes 0..15: one row each)
    auto issue = [&](int u, int st) {
        int s, tile, chunk;
        decode(u, s, tile, chunk);
        if (p.epi == E2_SWIGLU && chunk >= p.chunks) { s = 1; chunk -= p.chunks; }
        const G2Seg& sg = p.seg[s];
        const int cb = g2_chunk_blocks(sg.type), bb = type_block_bytes(sg.type), be = type_block_elems(sg.type);
        const int nb_row = K / be;
        const int nblk = min(cb, nb_row - chunk * cb);
        const uint32_t off = (uint32_t)(chunk * cb * bb);
        const uint32_t doff = off & 15u;
        const uint32_t bytes = (doff + (uint32_t)(nblk * bb) + 15u) & ~15u;
        const uint32_t bar = wbar + 8 * st;
        if (lane == 0) mbar_expect_tx(bar, 16u * bytes);
        __syncwarp();
        if (lane < 16) {
            const int row = min(tile * 16 + lane, sg.n_rows - 1);
            const uint8_t* src = sg.w + eoff * sg.expert_stride + (long long)row * sg.row_bytes + (off - doff);
            bulk_g2s(ring + (uint32_t)st * p.stage_bytes + (uint32_t)lane * g2_row_stride(sg.type), src, bytes, bar);
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
    if (threadIdx.x == 0) {
        mbar_expect_tx(xbar, xbytes);
        bulk_g2s(sm.xh, p.x.hi, 2u * K, xbar);
        bulk_g2s(sm.xl, p.x.lo, 2u * K, xbar);
        bulk_g2s(sm.xs, p.x.xs16, (uint32_t)(K >> 2), xbar);
    }
    if (warp == 0) {
        float inv = 1.0f;
        if (p.use_norm) {
            float ss = 0.0f;
            for (int i = lane; i < (K >> 4); i += 32) ss += p.x.ssq[i];
            ss = warp_sum(ss);
            inv = 1.0f / sqrtf(ss / (float)K + p.eps);
        }
        if (lane == 0) s_inv = inv;
    }
    __syncthreads();
    const float inv = s_inv;
    bool ok = mbar_wait(xbar, 0, p.err);

    float ag0 = 0.f, ag1 = 0.f, au0 = 0.f, au1 = 0.f;
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
        if (p.epi == E2_SWIGLU && chunk >= p.chunks) { is_up = true; ws = 1; chunk -= p.chunks; }
        const int type = p.seg[ws].type;
        const int cb = g2_chunk_blocks(type), be = type_block_elems(type);
        const int nblk = min(cb, K / be - chunk * cb);
        const int e0 = chunk * kG2ChunkElems;
        const uint32_t sp = ring + (uint32_t)st * p.stage_bytes;
        float a0 = 0.f, a1 = 0.f;
        switch (type) {
            case T_Q4_K: unit_k45<false>(sp, nblk, e0, sm, g, t, a0, a1); break;
            case T_Q5_K: unit_k45<true>(sp, nblk, e0, sm, g, t, a0, a1); break;
            case T_Q6_K: unit_q6k(sp, nblk, e0, (chunk * cb * 210) & 15, sm, g, t, a0, a1); break;
            default: unit_q80(sp, nblk, e0, sm, g, t, a0, a1); break;
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
        const int w_first = g2_owner(tu0, U, W), w_last = g2_owner(tu0 + p.units_per_tile - 1, U, W);
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
        const G2Seg& sg = p.seg[s];
        const int j = tile * 16 + (lane & 15);
        const bool valid = (lane < 16) && (j < sg.n_rows);
        v *= inv;
        float val = v;
        if (p.epi == E2_SWIGLU) {
            const float up = __shfl_sync(0xffffffffu, v, (lane & 15) + 16);
            val = g2_silu(v) * up;
        }
        if (valid && sg.bias) val += sg.bias[j];
        if (p.epi == E2_RESIDUAL && valid) val += p.residual[j];
        if (p.epi == E2_SCALED_ACC && valid) {  // moe.rs:363-368
            const float prev = p.expert_slot == 0 ? 0.0f : sg.out[j];
            val = prev + p.expert_wt[p.expert_slot] * val;
            if (p.residual) val += p.residual[j];
        }
        if (p.epi == E2_QKV_ROPE) {
            // Backend::rope, normal style: pairs (2i, 2i+1) sit in adjacent lanes (cpu/ops.rs:1322-1333)
            const float other = __shfl_xor_sync(0xffffffffu, val, 1);
            if (valid) {
                const int pos = *p.pos;
                const int head = j / p.hd, d = j - head * p.hd;
                if (s < 2) {
                    const float theta = ((float)pos / p.rope_scale) * p.freq[d >> 1];
                    const float c = cosf(theta), sn = sinf(theta);
                    const float r = (d & 1) ? __fadd_rn(__fmul_rn(other, sn), __fmul_rn(val, c))
                                            : __fsub_rn(__fmul_rn(val, c), __fmul_rn(other, sn));
                    if (s == 0) sg.out[j] = r;
                    else p.k_cache[((size_t)head * p.max_seq + pos) * p.hd + d] = r;
                } else {
                    p.v_cache[((size_t)head * p.max_seq + pos) * p.hd + d] = val;
                }
            }
        } else if (valid && sg.out) {
            sg.out[j] = val;
        }
        if (p.y.hi && s == 0) {  // prepared input of the next GEMV (uniform branch)
            float yv = valid ? val * (p.y_w ? p.y_w[j] : 1.0f) : 0.0f;
            yv = fminf(fmaxf(yv, -65504.0f), 65504.0f);
            const __half hh = __float2half_rn(yv);
            const __half ll = __float2half_rn(yv - __half2float(hh));
            float xs = __half2float(hh) + __half2float(ll);
            float sq = valid ? val * val : 0.0f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                xs += __shfl_xor_sync(0xffffffffu, xs, o);
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            if (valid) {
                p.y.hi[xperm(j)] = hh;
                p.y.lo[xperm(j)] = ll;
            }
            if (lane == 0) {
                p.y.xs16[tile] = xs;
                if (p.y.ssq) p.y.ssq[tile] = sq;
            }
        }
    }
}

}  // namespace b200
