// units2.cuh — low-register forms of the dequant-GEMV unit kernels of gemv_mma.cuh, for the second streamed megakernel
// (stream2.cuh: 14 consumer warps per SM at <= 128 registers instead of 7 at 255).
//
// Same arithmetic, same lane maps (LaneB, XSmem) and the same accumulation order per output row as unit_k45 / unit_q6k /
// unit_q80; what changes is the loop nest.  The first generation loads the B operands (the int8 planes of x) of a whole
// 256-element unit up front (32 registers) and unrolls both 16-row blocks: ~200 live registers.  Here the row-block loop is a
// run-time loop (its bounds let a job compute one 16-row block only) and the lane tables are three words (LaneT), which
// brings the same unit to ~80 live registers at the same ~270 instructions.  Measured dead ends of this round, kept for
// the record: splitting a unit between two warps by K halves (343 instructions per unit: scale decoding duplicated) or by
// row blocks (388: B operand loads duplicated); fully unrolled Q6_K variants (6 copies, 40 KB of hot code: instruction-cache bound).
// Reference arithmetic: src/backend/cpu/simd.rs:978-1146 (dot_q4_k / dot_q5_k / dot_q6_k / dot_q8_0).
#pragma once
#include "gemv_mma.cuh"

// the row-block loop of the unit4_* kernels: a run-time loop (<= 144 registers) or unrolled (8 fat consumer warps, stream2.cuh)
#if defined(B200_S2_CONS) && B200_S2_CONS == 8
#define S2_RT_UNROLL _Pragma("unroll")
#else
#define S2_RT_UNROLL _Pragma("unroll 1")
#endif

namespace b200 {

// Where a lane finds its B-operand bytes, in three registers (gemv_mma.cuh's LaneB holds the same information in eleven):
// c1 / c2 = shared address of the lane's first element in the (hi | mid) / lo plane at element 0; mk = which of the
// lane's B vectors are live (the others read the zero page).  Everything else (selectors, 1/16 for high nibbles) is
// recomputed from t inside the kernels: a handful of ALU instructions per half unit for eight registers.
struct LaneT {
    uint32_t c1, c2, mk;
    uint32_t grow;   // which of the 8 rows of a row block lane group g reads (bank-conflict-free weight loads, s2_row_perm)
};
// The 16-byte weight loads of a quarter-warp (lane groups 2j, 2j + 1: two rows x 64 bytes) are conflict-free when the two rows lie
// 64 bytes apart modulo 128.  The TMA box pitch is fixed by the GGUF block size (Q4_K: 144 or 288 bytes = 16 / 32 mod 128), so the
// ROW a lane group reads is permuted instead: a lane group may compute any row as long as its sums are filed under that row.
//   pitch = 32 mod 128 (two-chunk Q4_K entries): rows {0,2,1,3,4,6,5,7};  pitch = 16 mod 128 (Q4_K 144, Q8_0 272): rows {0,4,1,5,2,6,3,7}
__host__ __device__ inline uint32_t s2_row_perm(uint32_t pitch, int g) {
    const uint32_t m = pitch & 127u;
    if (m == 32u || m == 96u) return (uint32_t)((g & 4) | ((g & 1) << 1) | ((g >> 1) & 1));
    if (m == 16u || m == 48u || m == 80u || m == 112u) return (uint32_t)(((g & 1) << 2) | (g >> 1));
    return (uint32_t)g;
}
// lane group (0..7) that computed row r (0..7) of a row block
__host__ __device__ inline int s2_row_perm_inv(uint32_t pitch, int r) {
    for (int g = 0; g < 8; g++)
        if (s2_row_perm(pitch, g) == (uint32_t)r) return g;
    return r;
}
// Q4_K / Q5_K (lane_b_k45): lane (n, t) feeds column n with the k-slots of A-lane t; live only in the columns of the
// sub-block those k-slots belong to.  mk bits: 0 = low nibbles (hi | mid), 1 = high nibbles (hi | mid), 2 / 3 = the same for lo.
__device__ __forceinline__ LaneT lane_t_k45(uint32_t p0, uint32_t p1, uint32_t p2, int n, int t) {
    const bool act = (n >> 2) == (t >> 1);
    const int kind = (n >> 1) & 1;
    const uint32_t eo = 64u * (uint32_t)(t >> 1) + 16u * (uint32_t)(t & 1) + 32u * (uint32_t)kind;
    LaneT l;
    l.grow = (uint32_t)n;
    l.c1 = ((n & 1) ? p1 : p0) + eo;
    l.c2 = p2 + eo;
    const uint32_t m0 = (act && kind == 0) ? 1u : 0u, m1 = (act && kind == 1) ? 1u : 0u;
    l.mk = m0 | (m1 << 1) | (((n & 1) ? 0u : m0) << 2) | (((n & 1) ? 0u : m1) << 3);
    return l;
}
// Q6_K (lane_b_q6k): mk bits 0 / 1 = even / odd quarters of (hi | mid), bits 2..5 = quarter q of lo
__device__ __forceinline__ LaneT lane_t_q6k(uint32_t p0, uint32_t p1, uint32_t p2, int n, int t) {
    const bool act = ((n >> 1) & 1) == (t >> 1);
    LaneT l;
    l.grow = (uint32_t)n;
    l.c1 = ((n & 1) ? p1 : p0) + 8u * (uint32_t)t;
    l.c2 = p2 + 8u * (uint32_t)t;
    const int q2 = (n >> 2) + 2 * (n & 1);
    l.mk = act ? ((1u << (n >> 2)) | (4u << q2)) : 0u;
    return l;
}
// Q8_0 (lane_b_q80): mk = block (mod 4) the lane feeds in A1 | block it feeds in A2 << 4
__device__ __forceinline__ LaneT lane_t_q80(uint32_t p0, uint32_t p1, uint32_t p2, int n, int t) {
    LaneT l;
    l.grow = (uint32_t)n;
    l.c1 = ((n & 1) ? p1 : p0) + 8u * (uint32_t)t;
    l.c2 = p2 + 8u * (uint32_t)t;
    l.mk = (uint32_t)(n >> 1) | ((uint32_t)((n >> 1) + 4 * (n & 1)) << 4);
    return l;
}
// shared addresses of the per-32 scales / sums (sx), the per-16 Q6_K offsets (x16) and the zero page
struct XAddr {
    uint32_t sx, x16, zero;
};

// ---- Q4_K / Q5_K (blocks.rs:114-141): 32 rows x 256 elements, row blocks rt0 .. rt1 - 1 (a half job computes one of the two).
// acc[2 rt], acc[2 rt + 1] collect rows 16 rt + n and 16 rt + n + 8 (partial over the four t lanes).
template <bool Q5>
__device__ __forceinline__ void unit4_k45(uint32_t sp, uint32_t RS, uint32_t e0, const XAddr& sm, const LaneT& lt, int g, int t, int rt0,
                                          int rt1, float (&acc)[4]) {
    constexpr uint32_t QS = Q5 ? 48u : 16u;
    const uint32_t b1 = lt.c1 + e0, b2 = lt.c2 + e0;
    uint4 bl[2], bh[2], cl[2], ch[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        bl[c] = lds128(((lt.mk & 1u) ? b1 : sm.zero) + 128u * c);
        bh[c] = lds128(((lt.mk & 2u) ? b1 : sm.zero) + 128u * c);
        cl[c] = lds128(((lt.mk & 4u) ? b2 : sm.zero) + 128u * c);
        ch[c] = lds128(((lt.mk & 8u) ? b2 : sm.zero) + 128u * c);
    }
    const uint2 kx0 = lds64(sm.sx + 8u * ((e0 >> 5) + (uint32_t)t)), kx1 = lds64(sm.sx + 8u * ((e0 >> 5) + (uint32_t)t + 4u));
    const float hs = (!Q5 && (t & 1)) ? 0.0625f : 1.0f;   // high nibbles are carried x16
    const float kf[2] = {__uint_as_float(kx0.x) * hs, __uint_as_float(kx1.x) * hs};
    const float XS[2] = {__uint_as_float(kx0.y), __uint_as_float(kx1.y)};
    const uint32_t sel_yz = (uint32_t)t | ((uint32_t)(4 + t) << 4);
    S2_RT_UNROLL
    for (int rt = rt0; rt < rt1; rt++) {
        const uint32_t r0 = sp + ((uint32_t)(16 * rt) + lt.grow) * RS, r1 = r0 + 8u * RS;
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
                        hi[r] = src[r] & 0xF0F0F0F0u;   // 16 q: the 1/16 is in hs
                    }
                } else {   // 5th bit (dequant.rs:262-315): bit 2gp of the qh byte -> low sub-block, bit 2gp + 1 -> high sub-block
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
        // get_scale_min_k4 (dequant.rs:213-225) for sub-blocks t and 4 + t of the two rows: d * sc and dmin * m, one rounding each
        const uint32_t Y = __byte_perm(h0.y, h1.y, sel_yz), Z = __byte_perm(h0.z, h1.z, sel_yz);
        const uint32_t YZ = __byte_perm(Y, Z, 0x5410);              // (y r0, y r1, z r0, z r1): byte t of scales[0..3], [4..7]
        const uint32_t W = __byte_perm(h0.w, h1.w, sel_yz | (sel_yz << 8));   // (w r0, w r1, w r0, w r1): byte t of scales[8..11]
        const uint32_t W2 = (W & 0x00000F0Fu) | ((W >> 4) & 0x0F0F0000u);
        const uint32_t Rs[2] = {YZ & 0x3F3F3F3Fu, W2 | ((YZ >> 2) & 0x30303030u)};   // (sc r0, sc r1, m r0, m r1) of sub-block t / 4 + t
        const float d0 = half_bits_to_float(h0.x), n0 = half_bits_to_float(h0.x >> 16);
        const float d1 = half_bits_to_float(h1.x), n1 = half_bits_to_float(h1.x >> 16);
        const float d0b = d0 * -8388608.0f, n0b = n0 * -8388608.0f, d1b = d1 * -8388608.0f, n1b = n1 * -8388608.0f;
        float ra = 0.f, rb = 0.f;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const float dsc0 = fmaf(d0, byte_magic(Rs[c], 0x7440), d0b), dsc1 = fmaf(d1, byte_magic(Rs[c], 0x7441), d1b);
            const float dm0 = fmaf(n0, byte_magic(Rs[c], 0x7442), n0b), dm1 = fmaf(n1, byte_magic(Rs[c], 0x7443), n1b);
            const float s0 = fmaf((float)(A1[c][0] * 256 + A1[c][1]), 256.0f, (float)A2[c][0]);
            const float s1 = fmaf((float)(A1[c][2] * 256 + A1[c][3]), 256.0f, (float)A2[c][2]);
            ra = fmaf(dsc0 * kf[c], s0, fmaf(-dm0, XS[c], ra));
            rb = fmaf(dsc1 * kf[c], s1, fmaf(-dm1, XS[c], rb));
        }
        if (rt == 0) { acc[0] += ra; acc[1] += rb; } else { acc[2] += ra; acc[3] += rb; }
    }
}

// ---- Q6_K (blocks.rs:143-155, dequant.rs:321-356): as unit_q6k of gemv_mma.cuh, row blocks rt0 .. rt1 - 1
template <int AL>
__device__ __forceinline__ void unit4_q6k(uint32_t sp, uint32_t RS, uint32_t e0, const XAddr& sm, const LaneT& lt, int g, int t, int rt0,
                                          int rt1, float (&acc)[4]) {
    const uint32_t b1 = lt.c1 + e0, b2 = lt.c2 + e0;
    uint2 B1[2][4], B2[2][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t a1 = ((lt.mk >> (q & 1)) & 1u) ? b1 : sm.zero, a2 = ((lt.mk >> (2 + q)) & 1u) ? b2 : sm.zero;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            B1[hf][q] = lds64(a1 + 128u * hf + 32u * q);
            B2[hf][q] = lds64(a2 + 128u * hf + 32u * q);
        }
    }
#pragma unroll 1   // (unrolling the Q6_K / Q8_0 row-block loop doubles 20 KB of hot code: instruction-cache bound)
    for (int rt = rt0; rt < rt1; rt++) {
        const uint32_t blk0 = sp + ((uint32_t)(16 * rt) + lt.grow) * RS, blk1 = blk0 + 8u * RS;
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
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t a[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (q == 0) a[k] = (QA[k] & 0x0F0F0F0Fu) | ((QH[k] << 4) & 0x30303030u);
                    else if (q == 1) a[k] = (QB[k] & 0x0F0F0F0Fu) | ((QH[k] << 2) & 0x30303030u);
                    else if (q == 2) a[k] = ((QA[k] >> 4) & 0x0F0F0F0Fu) | (QH[k] & 0x30303030u);
                    else a[k] = ((QB[k] >> 4) & 0x0F0F0F0Fu) | ((QH[k] >> 2) & 0x30303030u);
                }
                if ((q & 1) == 0) imma_u8s8_z(A1[hf][q >> 1], a[0], a[1], a[2], a[3], B1[hf][q].x, B1[hf][q].y);
                else imma_u8s8(A1[hf][q >> 1], a[0], a[1], a[2], a[3], B1[hf][q].x, B1[hf][q].y);
                if (q == 0) imma_u8s8_z(A2[hf], a[0], a[1], a[2], a[3], B2[hf][q].x, B2[hf][q].y);
                else imma_u8s8(A2[hf], a[0], a[1], a[2], a[3], B2[hf][q].x, B2[hf][q].y);
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

// ---- Q8_0 (blocks.rs:60-70): as unit_q80 of gemv_mma.cuh, row blocks rt0 .. rt1 - 1
__device__ __forceinline__ void unit4_q80(uint32_t sp, uint32_t RS, uint32_t e0, int nblk, const XAddr& sm, const LaneT& lt, int g, int t,
                                          int rt0, int rt1, float (&acc)[4]) {
    const uint32_t t1 = lt.c1 + e0, t2 = lt.c2 + e0;
    uint2 B1[8], B2[8];
#pragma unroll
    for (int b = 0; b < 8; b++) {
        B1[b] = lds64(((uint32_t)(b & 3) == (lt.mk & 15u)) ? t1 + 32u * b : sm.zero);
        B2[b] = lds64(((uint32_t)b == (lt.mk >> 4)) ? t2 + 32u * b : sm.zero);
    }
#pragma unroll 1   // (unrolling the Q6_K / Q8_0 row-block loop doubles 20 KB of hot code: instruction-cache bound)
    for (int rt = rt0; rt < rt1; rt++) {
        const uint32_t row0 = sp + ((uint32_t)(16 * rt) + lt.grow) * RS, row1 = row0 + 8u * RS;
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

// ================= K-half outer / row-block inner forms (unit5_*): the B operands of ONE 128-element half live at a time (16 registers,
// shared by both row blocks), the scale pair of the half decoded per row block from re-read block headers.  ~300 instructions per unit
// at ~60 live registers: fits 14 consumer warps at 128 registers.
template <bool Q5>
__device__ __forceinline__ void unit5_k45(uint32_t sp, uint32_t RS, uint32_t e0, const XAddr& sm, const LaneT& lt, int g, int t, int rt0,
                                          int rt1, float (&acc)[4]) {
    constexpr uint32_t QS = Q5 ? 48u : 16u;
    const uint32_t sel_yz = (uint32_t)t | ((uint32_t)(4 + t) << 4);
    const float hs = (!Q5 && (t & 1)) ? 0.0625f : 1.0f;   // high nibbles are carried x16
#pragma unroll
    for (int C = 0; C < 2; C++) {
        const uint32_t b1 = lt.c1 + e0 + 128u * C, b2 = lt.c2 + e0 + 128u * C, z = sm.zero + 128u * C;
        const uint4 bl = lds128((lt.mk & 1u) ? b1 : z), bh = lds128((lt.mk & 2u) ? b1 : z);
        const uint4 cl = lds128((lt.mk & 4u) ? b2 : z), ch = lds128((lt.mk & 8u) ? b2 : z);
        const uint2 kx = lds64(sm.sx + 8u * ((e0 >> 5) + (uint32_t)t + 4u * C));
        const float kf = __uint_as_float(kx.x) * hs, XS = __uint_as_float(kx.y);
        const uint32_t bls[4] = {bl.x, bl.y, bl.z, bl.w}, bhs[4] = {bh.x, bh.y, bh.z, bh.w};
        const uint32_t cls[4] = {cl.x, cl.y, cl.z, cl.w}, chs[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll 1
        for (int rt = rt0; rt < rt1; rt++) {
            const uint32_t r0 = sp + ((uint32_t)(16 * rt) + lt.grow) * RS, r1 = r0 + 8u * RS;
            const uint4 h0 = lds128(r0), h1 = lds128(r1);
            const uint4 W0 = lds128(r0 + QS + 64u * C + 16u * (uint32_t)t), W1 = lds128(r1 + QS + 64u * C + 16u * (uint32_t)t);
            uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = qa;
            if (Q5) {
                qa = lds128(r0 + 16u + 16u * (uint32_t)(t & 1));
                qb = lds128(r1 + 16u + 16u * (uint32_t)(t & 1));
            }
            const uint32_t w0[4] = {W0.x, W0.y, W0.z, W0.w}, w1[4] = {W1.x, W1.y, W1.z, W1.w};
            const uint32_t ha[4] = {qa.x, qa.y, qa.z, qa.w}, hb[4] = {qb.x, qb.y, qb.z, qb.w};
            int A1[4], A2[4];
#pragma unroll
            for (int ip = 0; ip < 2; ip++) {
                uint32_t lo[4], hi[4];
                const uint32_t src[4] = {w0[2 * ip], w1[2 * ip], w0[2 * ip + 1], w1[2 * ip + 1]};   // fragment order a0..a3
                if (!Q5) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        lo[r] = src[r] & 0x0F0F0F0Fu;
                        hi[r] = src[r] & 0xF0F0F0F0u;   // 16 q: the 1/16 is in hs
                    }
                } else {   // 5th bit (dequant.rs:262-315): bit 2gp of the qh byte -> low sub-block, bit 2gp + 1 -> high sub-block
                    const uint32_t sh = 2u * (2u * C + (uint32_t)(t >> 1));
                    const uint32_t hq[4] = {ha[2 * ip] >> sh, hb[2 * ip] >> sh, ha[2 * ip + 1] >> sh, hb[2 * ip + 1] >> sh};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        lo[r] = (src[r] & 0x0F0F0F0Fu) | ((hq[r] << 4) & 0x10101010u);
                        hi[r] = ((src[r] >> 4) & 0x0F0F0F0Fu) | ((hq[r] << 3) & 0x10101010u);
                    }
                }
                if (ip == 0) {
                    imma_u8s8_z(A1, lo[0], lo[1], lo[2], lo[3], bls[0], bls[1]);
                    imma_u8s8_z(A2, lo[0], lo[1], lo[2], lo[3], cls[0], cls[1]);
                } else {
                    imma_u8s8(A1, lo[0], lo[1], lo[2], lo[3], bls[2], bls[3]);
                    imma_u8s8(A2, lo[0], lo[1], lo[2], lo[3], cls[2], cls[3]);
                }
                imma_u8s8(A1, hi[0], hi[1], hi[2], hi[3], bhs[2 * ip], bhs[2 * ip + 1]);
                imma_u8s8(A2, hi[0], hi[1], hi[2], hi[3], chs[2 * ip], chs[2 * ip + 1]);
            }
            // get_scale_min_k4 (dequant.rs:213-225) for sub-block t (C = 0) / 4 + t (C = 1) of the two rows
            const uint32_t Y = __byte_perm(h0.y, h1.y, sel_yz), Z = __byte_perm(h0.z, h1.z, sel_yz);
            const uint32_t YZ = __byte_perm(Y, Z, 0x5410);              // (y r0, y r1, z r0, z r1): byte t of scales[0..3], [4..7]
            uint32_t R;
            if (C == 0) {
                R = YZ & 0x3F3F3F3Fu;                                    // (sc_t r0, sc_t r1, m_t r0, m_t r1)
            } else {
                const uint32_t W = __byte_perm(h0.w, h1.w, sel_yz | (sel_yz << 8));   // (w r0, w r1, w r0, w r1): byte t of scales[8..11]
                R = (W & 0x00000F0Fu) | ((W >> 4) & 0x0F0F0000u) | ((YZ >> 2) & 0x30303030u);
            }
            const float d0 = half_bits_to_float(h0.x), n0 = half_bits_to_float(h0.x >> 16);
            const float d1 = half_bits_to_float(h1.x), n1 = half_bits_to_float(h1.x >> 16);
            const float dsc0 = fmaf(d0, byte_magic(R, 0x7440), d0 * -8388608.0f), dsc1 = fmaf(d1, byte_magic(R, 0x7441), d1 * -8388608.0f);
            const float dm0 = fmaf(n0, byte_magic(R, 0x7442), n0 * -8388608.0f), dm1 = fmaf(n1, byte_magic(R, 0x7443), n1 * -8388608.0f);
            const float s0 = fmaf((float)(A1[0] * 256 + A1[1]), 256.0f, (float)A2[0]);
            const float s1 = fmaf((float)(A1[2] * 256 + A1[3]), 256.0f, (float)A2[2]);
            if (rt == 0) {
                acc[0] = fmaf(dsc0 * kf, s0, fmaf(-dm0, XS, acc[0]));
                acc[1] = fmaf(dsc1 * kf, s1, fmaf(-dm1, XS, acc[1]));
            } else {
                acc[2] = fmaf(dsc0 * kf, s0, fmaf(-dm0, XS, acc[2]));
                acc[3] = fmaf(dsc1 * kf, s1, fmaf(-dm1, XS, acc[3]));
            }
        }
    }
}

template <int AL>
__device__ __forceinline__ void unit5_q6k(uint32_t sp, uint32_t RS, uint32_t e0, const XAddr& sm, const LaneT& lt, int g, int t, int rt0,
                                          int rt1, float (&acc)[4]) {
#pragma unroll 1
    for (uint32_t hf = 0; hf < 2; hf++) {
        const uint32_t b1 = lt.c1 + e0 + 128u * hf, b2 = lt.c2 + e0 + 128u * hf, z = sm.zero + 128u * hf;
        uint2 B1[4], B2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            B1[q] = lds64((((lt.mk >> (q & 1)) & 1u) ? b1 : z) + 32u * q);
            B2[q] = lds64((((lt.mk >> (2 + q)) & 1u) ? b2 : z) + 32u * q);
        }
#pragma unroll 1
        for (int rt = rt0; rt < rt1; rt++) {
            const uint32_t blk0 = sp + ((uint32_t)(16 * rt) + lt.grow) * RS, blk1 = blk0 + 8u * RS;
            uint32_t QA[4], QB[4], QH[4];   // fragment order: (row n word 0, row n+8 word 0, row n word 1, row n+8 word 1)
            lds_piece8<AL>(blk0 + 64u * hf + 8u * (uint32_t)t, QA[0], QA[2]);
            lds_piece8<AL>(blk1 + 64u * hf + 8u * (uint32_t)t, QA[1], QA[3]);
            lds_piece8<AL>(blk0 + 64u * hf + 32u + 8u * (uint32_t)t, QB[0], QB[2]);
            lds_piece8<AL>(blk1 + 64u * hf + 32u + 8u * (uint32_t)t, QB[1], QB[3]);
            lds_piece8<AL>(blk0 + 128u + 32u * hf + 8u * (uint32_t)t, QH[0], QH[2]);
            lds_piece8<AL>(blk1 + 128u + 32u * hf + 8u * (uint32_t)t, QH[1], QH[3]);
            int A1[2][4], A2[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t a[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (q == 0) a[k] = (QA[k] & 0x0F0F0F0Fu) | ((QH[k] << 4) & 0x30303030u);
                    else if (q == 1) a[k] = (QB[k] & 0x0F0F0F0Fu) | ((QH[k] << 2) & 0x30303030u);
                    else if (q == 2) a[k] = ((QA[k] >> 4) & 0x0F0F0F0Fu) | (QH[k] & 0x30303030u);
                    else a[k] = ((QB[k] >> 4) & 0x0F0F0F0Fu) | ((QH[k] >> 2) & 0x30303030u);
                }
                if ((q & 1) == 0) imma_u8s8_z(A1[q >> 1], a[0], a[1], a[2], a[3], B1[q].x, B1[q].y);
                else imma_u8s8(A1[q >> 1], a[0], a[1], a[2], a[3], B1[q].x, B1[q].y);
                if (q == 0) imma_u8s8_z(A2, a[0], a[1], a[2], a[3], B2[q].x, B2[q].y);
                else imma_u8s8(A2, a[0], a[1], a[2], a[3], B2[q].x, B2[q].y);
            }
            const float d0 = half_bits_to_float(lds16(blk0 + 208u)), d1 = half_bits_to_float(lds16(blk1 + 208u));
            float r0 = 0.f, r1 = 0.f;
#pragma unroll
            for (int qq = 0; qq < 2; qq++) {
                const uint32_t sg = 8u * hf + 4u * qq + (uint32_t)t;
                const float s0 = fmaf((float)(A1[qq][0] * 256 + A1[qq][1]), 256.0f, (float)A2[qq]);
                const float s1 = fmaf((float)(A1[qq][2] * 256 + A1[qq][3]), 256.0f, (float)A2[2 + qq]);
                const float kf = lds_f32(sm.sx + 8u * ((e0 >> 5) + (sg >> 1)));
                const float xn = lds_f32(sm.x16 + 4u * ((e0 >> 4) + sg));   // -32 * sum(x) of the group
                const float sc0 = (float)lds_s8(blk0 + 192u + sg), sc1 = (float)lds_s8(blk1 + 192u + sg);
                r0 = fmaf(d0 * sc0, fmaf(kf, s0, xn), r0);
                r1 = fmaf(d1 * sc1, fmaf(kf, s1, xn), r1);
            }
            if (rt == 0) { acc[0] += r0; acc[1] += r1; } else { acc[2] += r0; acc[3] += r1; }
        }
    }
}

__device__ __forceinline__ void unit5_q80(uint32_t sp, uint32_t RS, uint32_t e0, int nblk, const XAddr& sm, const LaneT& lt, int g, int t,
                                          int rt0, int rt1, float (&acc)[4]) {
    const uint32_t t1 = lt.c1 + e0, t2 = lt.c2 + e0;
#pragma unroll 1
    for (int J = 0; J < 2; J++) {
        if (4 * J >= nblk) break;   // warp-uniform (ragged last chunk)
        uint2 B1[4], B2[4];
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            const int b = 4 * J + bb;
            B1[bb] = lds64(((uint32_t)bb == (lt.mk & 15u)) ? t1 + 32u * b : sm.zero);
            B2[bb] = lds64(((uint32_t)b == (lt.mk >> 4)) ? t2 + 32u * b : sm.zero);
        }
#pragma unroll 1
        for (int rt = rt0; rt < rt1; rt++) {
            const uint32_t row0 = sp + ((uint32_t)(16 * rt) + lt.grow) * RS, row1 = row0 + 8u * RS;
            int A1[4] = {0, 0, 0, 0}, A2[4] = {0, 0, 0, 0};
#pragma unroll
            for (int bb = 0; bb < 4; bb++) {
                const int b = 4 * J + bb;
                if (b < nblk) {
                    uint32_t a0, a1, a2, a3;
                    lds_piece8_any(row0 + 34u * b + 2u + 8u * (uint32_t)t, a0, a2);
                    lds_piece8_any(row1 + 34u * b + 2u + 8u * (uint32_t)t, a1, a3);
                    imma_s8s8(A1, a0, a1, a2, a3, B1[bb].x, B1[bb].y);
                    imma_s8s8(A2, a0, a1, a2, a3, B2[bb].x, B2[bb].y);
                }
            }
            const int b = t + 4 * J;   // D-lane t owns blocks t and 4 + t
            if (b < nblk) {
                const float s0 = fmaf((float)(A1[0] * 256 + A1[1]), 256.0f, (float)(J ? A2[1] : A2[0]));
                const float s1 = fmaf((float)(A1[2] * 256 + A1[3]), 256.0f, (float)(J ? A2[3] : A2[2]));
                const float kf = lds_f32(sm.sx + 8u * ((e0 >> 5) + (uint32_t)b));
                const float v0 = half_bits_to_float(lds16(row0 + 34u * b)) * kf, v1 = half_bits_to_float(lds16(row1 + 34u * b)) * kf;
                if (rt == 0) { acc[0] = fmaf(v0, s0, acc[0]); acc[1] = fmaf(v1, s1, acc[1]); }
                else { acc[2] = fmaf(v0, s0, acc[2]); acc[3] = fmaf(v1, s1, acc[3]); }
            }
        }
    }
}

// The four accumulators of a lane (rows n, n + 8, 16 + n, 24 + n; partial over the four t lanes of a row group) ->
// lane L holds the complete sum of row L.  Same additions as (a += xor 1; a += xor 2) per register, 13 instructions
// instead of 28: each exchange halves the number of registers a lane keeps.
// ginv = lane group that computed row (lane & 7) of a row block (s2_row_perm_inv; the identity without a row permutation)
__device__ __forceinline__ float rows32_from_acc(const float (&acc)[4], int lane, int ginv) {
    const bool b0 = lane & 1, b1 = lane & 2;
    float k0 = b0 ? acc[1] : acc[0], s0 = b0 ? acc[0] : acc[1];
    float k1 = b0 ? acc[3] : acc[2], s1 = b0 ? acc[2] : acc[3];
    k0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    float k = b1 ? k1 : k0;
    const float s = b1 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 2);   // lane (n, t) holds row 8 t + n
    return __shfl_sync(0xffffffffu, k, 4 * ginv + (lane >> 3));
}

}  // namespace b200
