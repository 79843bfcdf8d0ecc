// gemv.cuh — fused dequant-GEMV for batch-1 decode: y[j] = sum_k deq(W)[j,k] * x[k].
//
// Replaces the reference's one-thread-per-output kernels vec_mat_q4k/q8_0/q6k/q5k
// (src/backend/cuda/kernels.rs:443-509, 519-553, 600-658, 661-735) and the CPU hot
// loop fused_vecmat_dispatch -> simd::dot_q* (src/backend/cpu/ops.rs:1123-1191,
// src/backend/cpu/simd.rs:931-1146).  Weights stay in the GGUF super-block layout,
// untransposed: row j of W is K/bs contiguous blocks.
//
// One launch covers up to three weight matrices that share x (q|k|v, gate|up) and
// fuses what surrounds the GEMV on the decode path:
//   prologue : RMSNorm of x (simd.rs:847-899) when norm_w != nullptr
//   epilogue : +bias (layers.rs:68-74) | +residual (layers.rs:1202-1241) |
//              silu(gate)*up (simd.rs:598-649) | out += w_e * y (moe.rs:363-368)
//
// Work split: a warp owns R=4 consecutive output rows and walks K; x lives in
// shared memory (padded layout, common.cuh:xidx) and each x chunk is reused by the
// 4 rows.  All weight loads of a step are issued before any is consumed
// (8 x 16 B in flight per lane for Q4_K).  Arithmetic is the reference's separated
// form d*sc*sum(q*x) - dmin*m*sum(x) in f32 (simd.rs:1006-1013).
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

constexpr int kGemvWarps = 8;
constexpr int kGemvThreads = kGemvWarps * kWarp;
constexpr int kGemvR = 4;  // rows per warp task

enum : int { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_SCALED_ACC = 3 };

struct GemvSeg {
    const uint8_t* w;       // n_rows rows of row_bytes
    float* out;             // n_rows outputs
    const float* bias;      // optional
    long long row_bytes;
    long long expert_stride;  // MoE: bytes between experts (0 = dense)
    int type;
    int n_rows;
};

struct GemvParams {
    GemvSeg seg[3];
    int n_seg;
    int K;
    const float* x;         // [K]
    const float* norm_w;    // optional fused RMSNorm weight [K]
    float eps;
    int epi;
    const float* residual;  // EPI_RESIDUAL: out[j] = acc + residual[j]
    // MoE (expert-resident, no host round trip): the slot-th selected expert
    const int* expert_sel;     // device [top_k] or nullptr
    const float* expert_wt;    // device [top_k] (EPI_SCALED_ACC: out[j] += wt * acc)
    int expert_slot;
    // expert parallel (one context per GPU, experts [expert_base, expert_base + expert_count) resident here): a launch whose selected
    // expert lives on another GPU returns at once; EPI_SCALED_ACC writes out[j] = wt * acc (no accumulate, no residual): the
    // weighted output of ONE expert, combined across GPUs by ep_combine_kernel (misc.cuh).  expert_count == 0: off.
    int expert_base, expert_count;
};

__device__ __forceinline__ float u8f(uint32_t w, int k) { return (float)((w >> (8 * k)) & 0xFFu); }

// ---- Q4_K / Q5_K: 8 lanes per 256-element block, 16 qs bytes (32 elements) per lane ----
template <bool Q5>
__device__ __forceinline__ void rows_dot_k45(const uint8_t* const (&rp)[kGemvR], int nb, const float* xs, int lane,
                                             float (&acc)[kGemvR]) {
    constexpr int BB = Q5 ? 176 : 144;
    constexpr int QS = Q5 ? 48 : 16;
    const int bl = lane >> 3, p = lane & 7, g = p >> 1, h = p & 1, gg = g & 1;
    for (int b0 = 0; b0 < nb; b0 += 4) {
        const int b = b0 + bl;
        const bool ok = b < nb;
        const int bc = ok ? b : nb - 1;
        uint4 hd[kGemvR], q[kGemvR], qh[kGemvR];
#pragma unroll
        for (int r = 0; r < kGemvR; r++) {
            const uint8_t* bp = rp[r] + (size_t)bc * BB;
            hd[r] = ldg_stream_u4(bp);
            q[r] = ldg_stream_u4(bp + QS + 16 * p);
            if (Q5) qh[r] = ldg_stream_u4(bp + 16 + 16 * h);
        }
        const float* xp = xs + xidx(bc * 256 + 64 * g + 16 * h);
        float xl[16], xh[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 a = *reinterpret_cast<const float4*>(xp + 4 * i);
            float4 c = *reinterpret_cast<const float4*>(xp + 32 + 4 * i);
            xl[4 * i] = a.x; xl[4 * i + 1] = a.y; xl[4 * i + 2] = a.z; xl[4 * i + 3] = a.w;
            xh[4 * i] = c.x; xh[4 * i + 1] = c.y; xh[4 * i + 2] = c.z; xh[4 * i + 3] = c.w;
        }
        float sl = 0.0f, sh = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; i++) { sl += xl[i]; sh += xh[i]; }
#pragma unroll
        for (int r = 0; r < kGemvR; r++) {
            const float d = half_bits_to_float(hd[r].x);
            const float dmin = half_bits_to_float(hd[r].x >> 16);
            const uint32_t A = (hd[r].y >> (16 * gg)) & 0xFFFFu;
            const uint32_t B = (hd[r].z >> (16 * gg)) & 0xFFFFu;
            const uint32_t C = (hd[r].w >> (16 * gg)) & 0xFFFFu;
            uint32_t scp, mnp;  // (sc[2g], sc[2g+1]) and (m[2g], m[2g+1]) as byte pairs
            if (g < 2) {
                scp = A & 0x3F3Fu;
                mnp = B & 0x3F3Fu;
            } else {
                scp = (C & 0x0F0Fu) | ((A >> 2) & 0x3030u);
                mnp = ((C >> 4) & 0x0F0Fu) | ((B >> 2) & 0x3030u);
            }
            const float d1 = d * (float)(scp & 0xFFu), d2 = d * (float)(scp >> 8);
            const float m1 = dmin * (float)(mnp & 0xFFu), m2 = dmin * (float)(mnp >> 8);
            const uint32_t qw[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
            const uint32_t hw[4] = {qh[r].x, qh[r].y, qh[r].z, qh[r].w};
            float ql = 0.0f, qhh = 0.0f;
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                uint32_t lo = qw[wi] & 0x0F0F0F0Fu;
                uint32_t hi = (qw[wi] >> 4) & 0x0F0F0F0Fu;
                if (Q5) {
                    lo |= ((hw[wi] >> (2 * g)) & 0x01010101u) << 4;
                    hi |= ((hw[wi] >> (2 * g + 1)) & 0x01010101u) << 4;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    ql = fmaf(u8f(lo, k), xl[4 * wi + k], ql);
                    qhh = fmaf(u8f(hi, k), xh[4 * wi + k], qhh);
                }
            }
            const float v = (d1 * ql - m1 * sl) + (d2 * qhh - m2 * sh);
            acc[r] += ok ? v : 0.0f;
        }
    }
}

// ---- Q6_K: 8 lanes per 210-byte block (2-byte aligned: 16-bit loads), 32 elements per lane ----
__device__ __forceinline__ void rows_dot_q6k(const uint8_t* const (&rp)[kGemvR], int nb, const float* xs, int lane,
                                             float (&acc)[kGemvR]) {
    const int bl = lane >> 3, t = lane & 7, n = t >> 2, j = t & 3, is = j >> 1;
    for (int b0 = 0; b0 < nb; b0 += 4) {
        const int b = b0 + bl;
        const bool ok = b < nb;
        const int bc = ok ? b : nb - 1;
        uint32_t qa[kGemvR][2], qb[kGemvR][2], qhv[kGemvR][2], dd[kGemvR];
        int sc[kGemvR][4];
#pragma unroll
        for (int r = 0; r < kGemvR; r++) {
            const uint8_t* bp = rp[r] + (size_t)bc * 210;
            const uint8_t* pa = bp + 64 * n + 8 * j;
            const uint8_t* ph = bp + 128 + 32 * n + 8 * j;
#pragma unroll
            for (int w = 0; w < 2; w++) {
                qa[r][w] = ldg_u16(pa + 4 * w) | (ldg_u16(pa + 4 * w + 2) << 16);
                qb[r][w] = ldg_u16(pa + 32 + 4 * w) | (ldg_u16(pa + 32 + 4 * w + 2) << 16);
                qhv[r][w] = ldg_u16(ph + 4 * w) | (ldg_u16(ph + 4 * w + 2) << 16);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) sc[r][c] = ldg_s8(bp + 192 + 8 * n + is + 2 * c);
            dd[r] = ldg_u16(bp + 208);
        }
        // x runs: element 128n + 32c + 8j + i, c = 0..3, i = 0..7
        float xv[4][8], xsum[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float* xp = xs + xidx(bc * 256 + 128 * n + 32 * c + 8 * j);
            float4 a = *reinterpret_cast<const float4*>(xp);
            float4 e = *reinterpret_cast<const float4*>(xp + 4);
            xv[c][0] = a.x; xv[c][1] = a.y; xv[c][2] = a.z; xv[c][3] = a.w;
            xv[c][4] = e.x; xv[c][5] = e.y; xv[c][6] = e.z; xv[c][7] = e.w;
            xsum[c] = ((a.x + a.y) + (a.z + a.w)) + ((e.x + e.y) + (e.z + e.w));
        }
#pragma unroll
        for (int r = 0; r < kGemvR; r++) {
            const float d = half_bits_to_float(dd[r]);
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int w = 0; w < 2; w++) {
                const uint32_t a = qa[r][w], bq = qb[r][w], hq = qhv[r][w];
                const uint32_t q0 = (a & 0x0F0F0F0Fu) | ((hq << 4) & 0x30303030u);
                const uint32_t q1 = (bq & 0x0F0F0F0Fu) | ((hq << 2) & 0x30303030u);
                const uint32_t q2 = ((a >> 4) & 0x0F0F0F0Fu) | (hq & 0x30303030u);
                const uint32_t q3 = ((bq >> 4) & 0x0F0F0F0Fu) | ((hq >> 2) & 0x30303030u);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    s[0] = fmaf(u8f(q0, k), xv[0][4 * w + k], s[0]);
                    s[1] = fmaf(u8f(q1, k), xv[1][4 * w + k], s[1]);
                    s[2] = fmaf(u8f(q2, k), xv[2][4 * w + k], s[2]);
                    s[3] = fmaf(u8f(q3, k), xv[3][4 * w + k], s[3]);
                }
            }
            float v = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; c++) v += (d * (float)sc[r][c]) * (s[c] - 32.0f * xsum[c]);
            acc[r] += ok ? v : 0.0f;
        }
    }
}

// ---- 32-element block types (Q8_0, Q4_0, Q5_0) and F16/F32: one lane per 32 elements ----
__device__ __forceinline__ void rows_dot_b32(int type, const uint8_t* const (&rp)[kGemvR], int K, const float* xs,
                                             int lane, float (&acc)[kGemvR]) {
    const int nb = K >> 5;
    const int bb = (type == T_F32) ? 128 : (type == T_F16) ? 64 : type_block_bytes(type);
    for (int b = lane; b < nb; b += 32) {
        float xv[32];
        const float* xp = xs + xidx(b * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 a = *reinterpret_cast<const float4*>(xp + 4 * i);
            xv[4 * i] = a.x; xv[4 * i + 1] = a.y; xv[4 * i + 2] = a.z; xv[4 * i + 3] = a.w;
        }
#pragma unroll
        for (int r = 0; r < kGemvR; r++) {
            const uint8_t* bp = rp[r] + (size_t)b * bb;
            float s = 0.0f;
            if (type == T_Q8_0) {
                const float d = half_bits_to_float(ldg_u16(bp));
                uint32_t qv[16];
#pragma unroll
                for (int i = 0; i < 16; i++) qv[i] = ldg_u16(bp + 2 + 2 * i);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    s = fmaf((float)(int)(signed char)(qv[i] & 0xFFu), xv[2 * i], s);
                    s = fmaf((float)(int)(signed char)(qv[i] >> 8), xv[2 * i + 1], s);
                }
                s *= d;
            } else if (type == T_Q4_0) {
                const float d = half_bits_to_float(ldg_u16(bp));
                float lo = 0.0f, hi = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t v = ldg_u16(bp + 2 + 2 * i);
                    lo = fmaf((float)((int)(v & 0xF) - 8), xv[2 * i], lo);
                    hi = fmaf((float)((int)((v >> 4) & 0xF) - 8), xv[2 * i + 16], hi);
                    lo = fmaf((float)((int)((v >> 8) & 0xF) - 8), xv[2 * i + 1], lo);
                    hi = fmaf((float)((int)((v >> 12) & 0xF) - 8), xv[2 * i + 17], hi);
                }
                s = d * (lo + hi);
            } else if (type == T_Q5_0) {
                const float d = half_bits_to_float(ldg_u16(bp));
                const uint32_t qh = ldg_u16(bp + 2) | (ldg_u16(bp + 4) << 16);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t v = ldg_u16(bp + 6 + 2 * i);
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        int e = 2 * i + k;
                        int byte = (v >> (8 * k)) & 0xFF;
                        int lo = ((byte & 0xF) | (((qh >> e) & 1) << 4)) - 16;
                        int hi = ((byte >> 4) | (((qh >> (e + 16)) & 1) << 4)) - 16;
                        s = fmaf((float)lo, xv[e], s);
                        s = fmaf((float)hi, xv[e + 16], s);
                    }
                }
                s *= d;
            } else if (type == T_F16) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t v = ldg_stream_u32(bp + 4 * i);
                    s = fmaf(half_bits_to_float(v), xv[2 * i], s);
                    s = fmaf(half_bits_to_float(v >> 16), xv[2 * i + 1], s);
                }
            } else {  // F32
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint4 v = ldg_stream_u4(bp + 16 * i);
                    s = fmaf(__uint_as_float(v.x), xv[4 * i], s);
                    s = fmaf(__uint_as_float(v.y), xv[4 * i + 1], s);
                    s = fmaf(__uint_as_float(v.z), xv[4 * i + 2], s);
                    s = fmaf(__uint_as_float(v.w), xv[4 * i + 3], s);
                }
            }
            acc[r] += s;
        }
    }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// Stage x (optionally RMS-normalised) into shared memory in the padded layout.
__device__ __forceinline__ void stage_x(const GemvParams& p, float* xs, float* red) {
    const int tid = threadIdx.x;
    float inv = 1.0f;
    if (p.norm_w) {
        float ss = 0.0f;
        for (int e = tid * 4; e < p.K; e += kGemvThreads * 4) {
            float4 v = *reinterpret_cast<const float4*>(p.x + e);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        if ((tid & 31) == 0) red[tid >> 5] = ss;
        __syncthreads();
        float tot = 0.0f;
#pragma unroll
        for (int w = 0; w < kGemvWarps; w++) tot += red[w];
        inv = 1.0f / sqrtf(tot / (float)p.K + p.eps);
    }
    for (int e = tid * 4; e < p.K; e += kGemvThreads * 4) {
        float4 v = *reinterpret_cast<const float4*>(p.x + e);
        if (p.norm_w) {
            float4 w = *reinterpret_cast<const float4*>(p.norm_w + e);
            v.x = (v.x * inv) * w.x; v.y = (v.y * inv) * w.y; v.z = (v.z * inv) * w.z; v.w = (v.w * inv) * w.w;
        }
        *reinterpret_cast<float4*>(xs + xidx(e)) = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kGemvThreads) gemv_kernel(const GemvParams p) {
    extern __shared__ __align__(16) float xs[];
    __shared__ float red[kGemvWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // task table: segment s owns tasks [t0[s], t0[s+1]); a task = kGemvR rows
    // (EPI_SWIGLU: 2 gate rows + the 2 matching up rows).
    int t0[4];
    t0[0] = 0;
    if (p.epi == EPI_SWIGLU) {
        t0[1] = (p.seg[0].n_rows + 1) / 2;
        t0[2] = t0[3] = t0[1];
    } else {
#pragma unroll
        for (int s = 0; s < 3; s++) t0[s + 1] = t0[s] + (s < p.n_seg ? (p.seg[s].n_rows + kGemvR - 1) / kGemvR : 0);
    }
    const int n_tasks = t0[3];

    pdl_launch_dependents();
    if (!p.expert_sel) {  // dense weights do not depend on the predecessor: pull this warp's first rows towards L2
        int task = blockIdx.x * kGemvWarps + warp;
        if (task < n_tasks) {
            if (p.epi == EPI_SWIGLU) {
                const long long off = (long long)task * 2;
                const long long bytes = 2 * p.seg[0].row_bytes;
                for (long long o = (long long)lane * 128; o < bytes; o += 32 * 128) {
                    prefetch_l2(p.seg[0].w + off * p.seg[0].row_bytes + o);
                    prefetch_l2(p.seg[1].w + off * p.seg[1].row_bytes + o);
                }
            } else {
                int s = (task >= t0[2]) ? 2 : (task >= t0[1]) ? 1 : 0;
                const GemvSeg& sg = p.seg[s];
                int row0 = (task - t0[s]) * kGemvR;
                int rows = min(kGemvR, sg.n_rows - row0);
                const uint8_t* base = sg.w + (long long)row0 * sg.row_bytes;
                long long bytes = (long long)rows * sg.row_bytes;
                for (long long o = (long long)lane * 128; o < bytes; o += 32 * 128) prefetch_l2(base + o);
            }
        }
    }
    pdl_wait();
    // MoE: the selected expert is written by the routing kernel (a predecessor): read it only after the wait
    long long eoff = p.expert_sel ? (long long)p.expert_sel[p.expert_slot] : 0;
    if (p.expert_count > 0) {   // expert parallel: not this GPU's expert -> nothing to do (uniform over the grid)
        eoff -= p.expert_base;
        if (eoff < 0 || eoff >= p.expert_count) return;
    }
    stage_x(p, xs, red);

    for (int task = blockIdx.x * kGemvWarps + warp; task < n_tasks; task += gridDim.x * kGemvWarps) {
        const uint8_t* rp[kGemvR];
        int type, s = 0, row0, rows;
        if (p.epi == EPI_SWIGLU) {
            const GemvSeg& ga = p.seg[0];
            const GemvSeg& up = p.seg[1];
            row0 = task * 2;
            rows = min(2, ga.n_rows - row0);
            type = ga.type;
            const int r1 = row0 + (rows > 1 ? 1 : 0);
            rp[0] = ga.w + eoff * ga.expert_stride + (long long)row0 * ga.row_bytes;
            rp[1] = ga.w + eoff * ga.expert_stride + (long long)r1 * ga.row_bytes;
            rp[2] = up.w + eoff * up.expert_stride + (long long)row0 * up.row_bytes;
            rp[3] = up.w + eoff * up.expert_stride + (long long)r1 * up.row_bytes;
        } else {
            s = (task >= t0[2]) ? 2 : (task >= t0[1]) ? 1 : 0;
            const GemvSeg& sg = p.seg[s];
            row0 = (task - t0[s]) * kGemvR;
            rows = min(kGemvR, sg.n_rows - row0);
            type = sg.type;
#pragma unroll
            for (int r = 0; r < kGemvR; r++)
                rp[r] = sg.w + eoff * sg.expert_stride + (long long)(row0 + min(r, rows - 1)) * sg.row_bytes;
        }
        float acc[kGemvR] = {0.0f, 0.0f, 0.0f, 0.0f};
        switch (type) {
            case T_Q4_K: rows_dot_k45<false>(rp, p.K >> 8, xs, lane, acc); break;
            case T_Q5_K: rows_dot_k45<true>(rp, p.K >> 8, xs, lane, acc); break;
            case T_Q6_K: rows_dot_q6k(rp, p.K >> 8, xs, lane, acc); break;
            default: rows_dot_b32(type, rp, p.K, xs, lane, acc); break;
        }
#pragma unroll
        for (int r = 0; r < kGemvR; r++) acc[r] = warp_sum(acc[r]);
        if (lane == 0) {
            if (p.epi == EPI_SWIGLU) {
                float* out = p.seg[0].out;
                out[row0] = silu_f(acc[0]) * acc[2];
                if (rows > 1) out[row0 + 1] = silu_f(acc[1]) * acc[3];
            } else {
                const GemvSeg& sg = p.seg[s];
#pragma unroll
                for (int r = 0; r < kGemvR; r++) {
                    if (r < rows) {
                        const int jrow = row0 + r;
                        float v = acc[r];
                        if (sg.bias) v += sg.bias[jrow];
                        if (p.epi == EPI_RESIDUAL) v += p.residual[jrow];
                        if (p.epi == EPI_SCALED_ACC) {  // moe.rs:363-368: out (zeros) += w_e * y_e, in selection order
                            const float prev = (p.expert_slot == 0 || p.expert_count > 0) ? 0.0f : sg.out[jrow];
                            v = prev + p.expert_wt[p.expert_slot] * v;
                            if (p.residual && p.expert_count == 0) v += p.residual[jrow];  // last selected expert: + h (layers.rs:1235-1241)
                        }
                        sg.out[jrow] = v;
                    }
                }
            }
        }
    }
}

}  // namespace b200
