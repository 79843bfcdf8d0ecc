// units2.cuh — low-register forms of the dequant-GEMV unit kernels of gemv_mma.cuh, for the second streamed megakernel
// (stream2.cuh: 14 consumer warps per SM at <= 128 registers instead of 7 at 255).
//
// Same arithmetic, same lane maps (LaneB, XSmem) and the same accumulation order per output row as unit_k45 / unit_q6k /
// unit_q80; what changes is the loop nest.  The first generation loads the B operands (the int8 planes of x) of a whole
// 256-element unit up front (32 registers) and unrolls both 16-row blocks: ~200 live registers.  Here a kernel does ONE
// K half of a unit (Q4_K / Q5_K: 128-element half C, Q6_K: half HF, Q8_0: four blocks J): only that half's B operands
// (16 registers) and one row block's accumulators are live, block headers are re-read per half.  In stream2.cuh the two
// halves of a ring entry are computed by the two warps of a PAIR (even warp: half 0, odd warp: half 1), which halves
// the granularity work is dealt at and keeps the B-operand loads of the pair disjoint.
// Reference arithmetic: src/backend/cpu/simd.rs:978-1146 (dot_q4_k / dot_q5_k / dot_q6_k / dot_q8_0).
#pragma once
#include "gemv_mma.cuh"

namespace b200 {

// get_scale_min_k4 (dequant.rs:213-225) for ONE sub-block per row (t for C = 0, 4 + t for C = 1) of two rows
template <int C>
__device__ __forceinline__ void k4_scales_half(const uint4& h0, const uint4& h1, const LaneB& lb, float (&dsc)[2], float (&dm)[2]) {
    const uint32_t Y = __byte_perm(h0.y, h1.y, lb.sel_yz), Z = __byte_perm(h0.z, h1.z, lb.sel_yz);
    const uint32_t YZ = __byte_perm(Y, Z, 0x5410);              // (y r0, y r1, z r0, z r1): byte t of scales[0..3], [4..7]
    uint32_t R;
    if (C == 0) {
        R = YZ & 0x3F3F3F3Fu;                                    // (sc_t r0, sc_t r1, m_t r0, m_t r1)
    } else {
        const uint32_t W = __byte_perm(h0.w, h1.w, lb.sel_w);   // (w r0, w r1, w r0, w r1): byte t of scales[8..11]
        const uint32_t W2 = (W & 0x00000F0Fu) | ((W >> 4) & 0x0F0F0000u);
        R = W2 | ((YZ >> 2) & 0x30303030u);                      // (sc_{4+t} r0, r1, m_{4+t} r0, r1)
    }
    const float d0 = half_bits_to_float(h0.x), n0 = half_bits_to_float(h0.x >> 16);
    const float d1 = half_bits_to_float(h1.x), n1 = half_bits_to_float(h1.x >> 16);
    dsc[0] = fmaf(d0, byte_magic(R, 0x7440), d0 * -8388608.0f);
    dsc[1] = fmaf(d1, byte_magic(R, 0x7441), d1 * -8388608.0f);
    dm[0] = fmaf(n0, byte_magic(R, 0x7442), n0 * -8388608.0f);
    dm[1] = fmaf(n1, byte_magic(R, 0x7443), n1 * -8388608.0f);
}

template <bool Q5, int C>
__device__ __forceinline__ void unit2_k45_half(uint32_t sp, uint32_t RS, uint32_t e0, const XSmem& sm, const LaneB& lb, int g, int t,
                                               float (&acc)[4]) {
    constexpr uint32_t QS = Q5 ? 48u : 16u;
    const uint32_t t1 = lb.d1 + e0, t2 = lb.d2 + e0;
    const uint4 bl = lds128(sm.zero + lb.m[0] * t1 + 128u * C), bh = lds128(sm.zero + lb.m[1] * t1 + 128u * C);
    const uint4 cl = lds128(sm.zero + lb.m[2] * t2 + 128u * C), ch = lds128(sm.zero + lb.m[3] * t2 + 128u * C);
    const uint2 kx = lds64(sm.sx + 8u * ((e0 >> 5) + (uint32_t)t + 4u * C));
    const float kf = __uint_as_float(kx.x) * lb.hs, XS = __uint_as_float(kx.y);
    const uint32_t bls[4] = {bl.x, bl.y, bl.z, bl.w}, bhs[4] = {bh.x, bh.y, bh.z, bh.w};
    const uint32_t cls[4] = {cl.x, cl.y, cl.z, cl.w}, chs[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
    for (int rt = 0; rt < 2; rt++) {
        const uint32_t r0 = sp + (uint32_t)(16 * rt + g) * RS, r1 = r0 + 8u * RS;
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
                    hi[r] = src[r] & 0xF0F0F0F0u;   // 16 q: the 1/16 is in lb.hs
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
        float dsc[2], dm[2];
        k4_scales_half<C>(h0, h1, lb, dsc, dm);
        const float s0 = fmaf((float)(A1[0] * 256 + A1[1]), 256.0f, (float)A2[0]);
        const float s1 = fmaf((float)(A1[2] * 256 + A1[3]), 256.0f, (float)A2[2]);
        acc[2 * rt] = fmaf(dsc[0] * kf, s0, fmaf(-dm[0], XS, acc[2 * rt]));
        acc[2 * rt + 1] = fmaf(dsc[1] * kf, s1, fmaf(-dm[1], XS, acc[2 * rt + 1]));
    }
}
// Q6_K: one 128-element half HF (B operands of the half: 16 registers)
template <int AL, int HF>
__device__ __forceinline__ void unit2_q6k_half(uint32_t sp, uint32_t RS, uint32_t e0, const XSmem& sm, const LaneB& lb, int g, int t,
                                               float (&acc)[4]) {
    const uint32_t t1 = lb.d1 + e0, t2 = lb.d2 + e0;
    const uint32_t a1[2] = {sm.zero + lb.m[0] * t1, sm.zero + lb.m[1] * t1};
    uint2 B1[4], B2[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        B1[q] = lds64(a1[q & 1] + 128u * HF + 32u * q);
        B2[q] = lds64(sm.zero + lb.m[2 + q] * t2 + 128u * HF + 32u * q);
    }
#pragma unroll
    for (int rt = 0; rt < 2; rt++) {
        const uint32_t blk0 = sp + (uint32_t)(16 * rt + g) * RS, blk1 = blk0 + 8u * RS;
        uint32_t QA[4], QB[4], QH[4];   // fragment order: (row n word 0, row n+8 word 0, row n word 1, row n+8 word 1)
        lds_piece8<AL>(blk0 + 64u * HF + 8u * (uint32_t)t, QA[0], QA[2]);
        lds_piece8<AL>(blk1 + 64u * HF + 8u * (uint32_t)t, QA[1], QA[3]);
        lds_piece8<AL>(blk0 + 64u * HF + 32u + 8u * (uint32_t)t, QB[0], QB[2]);
        lds_piece8<AL>(blk1 + 64u * HF + 32u + 8u * (uint32_t)t, QB[1], QB[3]);
        lds_piece8<AL>(blk0 + 128u + 32u * HF + 8u * (uint32_t)t, QH[0], QH[2]);
        lds_piece8<AL>(blk1 + 128u + 32u * HF + 8u * (uint32_t)t, QH[1], QH[3]);
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
            const uint32_t sg = 8u * HF + 4u * qq + (uint32_t)t;
            const float s0 = fmaf((float)(A1[qq][0] * 256 + A1[qq][1]), 256.0f, (float)A2[qq]);
            const float s1 = fmaf((float)(A1[qq][2] * 256 + A1[qq][3]), 256.0f, (float)A2[2 + qq]);
            const float kf = lds_f32(sm.sx + 8u * ((e0 >> 5) + (sg >> 1)));
            const float xn = lds_f32(sm.x16 + 4u * ((e0 >> 4) + sg));   // -32 * sum(x) of the group
            const float sc0 = (float)lds_s8(blk0 + 192u + sg), sc1 = (float)lds_s8(blk1 + 192u + sg);
            r0 = fmaf(d0 * sc0, fmaf(kf, s0, xn), r0);
            r1 = fmaf(d1 * sc1, fmaf(kf, s1, xn), r1);
        }
        acc[2 * rt] += r0;
        acc[2 * rt + 1] += r1;
    }
}

// Q8_0: the four 34-byte blocks 4J .. 4J + 3 of a 256-element chunk
template <int J>
__device__ __forceinline__ void unit2_q80_half(uint32_t sp, uint32_t RS, uint32_t e0, int nblk, const XSmem& sm, const LaneB& lb, int g, int t,
                                               float (&acc)[4]) {
    if (4 * J >= nblk) return;   // warp-uniform (ragged last chunk)
    const uint32_t t1 = sm.zero + lb.d1 + e0, t2 = sm.zero + lb.d2 + e0;
    uint2 B1[4], B2[4];
#pragma unroll
    for (int bb = 0; bb < 4; bb++) {
        const int b = 4 * J + bb;
        B1[bb] = lds64(((uint32_t)bb == lb.m[0]) ? t1 + 32u * b : sm.zero);
        B2[bb] = lds64(((uint32_t)b == lb.m[1]) ? t2 + 32u * b : sm.zero);
    }
#pragma unroll
    for (int rt = 0; rt < 2; rt++) {
        const uint32_t row0 = sp + (uint32_t)(16 * rt + g) * RS, row1 = row0 + 8u * RS;
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
            const float s0 = fmaf((float)(A1[0] * 256 + A1[1]), 256.0f, (float)A2[J]);
            const float s1 = fmaf((float)(A1[2] * 256 + A1[3]), 256.0f, (float)A2[2 + J]);
            const float kf = lds_f32(sm.sx + 8u * ((e0 >> 5) + (uint32_t)b));
            acc[2 * rt] = fmaf(half_bits_to_float(lds16(row0 + 34u * b)) * kf, s0, acc[2 * rt]);
            acc[2 * rt + 1] = fmaf(half_bits_to_float(lds16(row1 + 34u * b)) * kf, s1, acc[2 * rt + 1]);
        }
    }
}

}  // namespace b200
