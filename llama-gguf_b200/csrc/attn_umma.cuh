// attn_umma.cuh — causal GQA attention of a prompt chunk on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Same function as prefill_attn_kernel (prefill.cuh) and as the reference's attention_cached applied to T rows
// (src/backend/cpu/ops.rs:1479-1537: s[p] = dot(q, K[p]) * scale, softmax over p <= pos, out = sum s[p] * V[p]; GQA map
// kv_head = head / (n_heads / n_kv), ops.rs:1513), but the two contractions run as tcgen05.mma instead of one warp per
// (token, kv head) on CUDA cores — that kernel was 40 % of a 2048-token prefill and had no tensor instruction at all.
//
// A CTA owns 128 query ROWS = (128 / G tokens) x (the G query heads of one kv head): the rows share every K / V tile.
//   prep   prefill_kv16_kernel rounds the f32 cache rows 0 .. kv_end-1 of the layer to fp16 once per chunk:
//          K16[kv][pos][hd] and the transpose Vt16[kv][hd][pos] (both operands of a tcgen05.mma are K-major), pad rows zeroed;
//   pass 1 per 128-key tile: TMA (SWIZZLE_128B boxes) -> S = Q K^T (hd/16 x tcgen05.mma kind::f16, M = N = 128) into TMEM,
//          thread r reads row r (tcgen05.ld), scales, masks keys > its query position, keeps the running max and sum;
//   pass 2 per tile: S again, e = exp(s - max) into a swizzled P tile, O += P V (tcgen05.mma, N = hd) accumulated in TMEM; at the
//          end O / sum -> fp16 output row.  P and V are carried as fp16 PAIRS (hi + lo, 22 significant bits) and the product as
//          P_hi V_hi + P_lo V_hi + P_hi V_lo: an attention output is an average with heavy cancellation, so single-fp16 P and V put
//          2^-12 of relative noise on it — measured: the logits of a 300-token prompt moved from 6e-4 to 1.1e-3 of the oracle's.
// Two passes instead of an online rescale of O: the softmax is the reference's (global max, then exp, then normalise) and
// nothing in TMEM is ever rewritten; the price is a second Q K^T, half of which the tensor pipe does while idle anyway.
// fp16 operands (q, k; v and e as hi + lo pairs), f32 scores / softmax / accumulation.  Every wait is bounded and reports through p.err.
#pragma once
#include <cuda.h>
#include "gemm_umma2.cuh"

namespace b200 {

struct AttnUmmaParams {
    const float* qkv;      // [T][ld], q rotated (f32)
    int ld;
    __half* out;           // [T][ldo]: [n_heads][hd] per token, fp16
    int ldo;
    int pos0, T, n_heads, n_kv, G;
    int P;                 // padded position capacity of the fp16 copies (multiple of 128)
    float scale;
    int* err;
};

// fp16 hi / lo pairs of one layer's cache rows 0 .. kv_end-1 (zero up to kv_pad): K16[kv][pos][hd] and the transpose
// Vt16[kv][hd][pos]; hi = fp16(x), lo = fp16(x - hi); the lo halves follow the hi halves (rows n_kv * P ... / n_kv * hd ...).
// grid (kv_pad / 64, n_kv), 256 threads; HD = 64 or 128.
template <int HD>
__global__ void __launch_bounds__(256) prefill_kv16_kernel(const float* __restrict__ kc, const float* __restrict__ vc, int max_seq, int kv_end, int P,
                                                           __half* __restrict__ k16, __half* __restrict__ vt16) {
    __shared__ __half tile[64][HD + 2], tile_lo[64][HD + 2];
    const int kh = blockIdx.y, n_kv = gridDim.y, p0 = blockIdx.x * 64;
    const float* ks = kc + (size_t)kh * max_seq * HD;
    const float* vs = vc + (size_t)kh * max_seq * HD;
    for (int i = threadIdx.x; i < 64 * HD; i += 256) {
        const int pr = i / HD, d = i - pr * HD, pos = p0 + pr;
        const bool ok = pos < kv_end;
        const float k = ok ? fminf(fmaxf(ks[(size_t)pos * HD + d], -65504.0f), 65504.0f) : 0.0f;
        const float v = ok ? fminf(fmaxf(vs[(size_t)pos * HD + d], -65504.0f), 65504.0f) : 0.0f;
        const __half khi = __float2half_rn(k), vhi = __float2half_rn(v);
        k16[((size_t)kh * P + pos) * HD + d] = khi;
        k16[((size_t)(n_kv + kh) * P + pos) * HD + d] = __float2half_rn(k - __half2float(khi));
        tile[pr][d] = vhi;
        tile_lo[pr][d] = __float2half_rn(v - __half2float(vhi));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * HD; i += 256) {
        const int d = i >> 6, pr = i & 63;
        vt16[((size_t)kh * HD + d) * P + p0 + pr] = tile[pr][d];
        vt16[((size_t)(n_kv + kh) * HD + d) * P + p0 + pr] = tile_lo[pr][d];
    }
}

constexpr int kAttnTK = 64;   // keys per tile

// 256 threads: threads t and t + 128 share query row t & 127 (TMEM lane t & 127) and split every tile's 64 key columns
// (and the head dimension in the prologue / epilogue); thread 0 issues the TMA loads and the MMAs.
template <int HD>
__global__ void __launch_bounds__(256, 1) prefill_attn_umma_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                                                                   const AttnUmmaParams p) {
    extern __shared__ __align__(1024) uint8_t au_smem[];
    __shared__ __align__(8) unsigned long long s_bar[4];   // k_full, v_full, s_done, pv_done
    __shared__ uint32_t s_tmem;
    __shared__ float s_ml[2][128][2];
    constexpr int NS = HD / 64;                     // 64-element slabs along hd
    constexpr int kQBytes = 128 * HD * 2;           // Q tile (hi or lo): NS slabs of [128 rows][128 B]
    constexpr int kKBytes = kAttnTK * HD * 2;       // K tile (hi or lo): NS slabs of [64 keys][128 B]
    constexpr int kVBytes = HD * kAttnTK * 2;       // Vt tile (hi or lo): [HD rows][64 keys = 128 B]
    constexpr int kPBytes = 128 * kAttnTK * 2;      // P tile (hi or lo): [128 rows][64 keys = 128 B]
    constexpr int kCols = 256;                      // S: columns 0..63, O: columns 128..128+HD-1
    uint8_t* base = au_smem + ((1024u - (umma_smem_u32(au_smem) & 1023u)) & 1023u);
    uint8_t* sQ = base;                             // hi, lo
    uint8_t* sK = sQ + 2 * kQBytes;                 // hi, lo
    uint8_t* sV = sK + 2 * kKBytes;                 // hi, lo
    uint8_t* sP = sV + 2 * kVBytes;                 // hi, lo
    const int tid = threadIdx.x, warp = tid >> 5, half = tid >> 7;
    const int G = p.G, TQ = 128 / G;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;   // heavy (late) query tiles first
    const int kh = blockIdx.y, t0 = qt * TQ;
    const uint32_t bar = umma_smem_u32(&s_bar[0]);
    const uint32_t bK = bar, bV = bar + 8u, bS = bar + 16u, bPV = bar + 24u;
    if (tid == 0) {
        for (int i = 0; i < 4; i++) u2_mbar_init(bar + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma_smem_u32(&s_tmem)), "n"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- Q tile: row r = (token t0 + r / G, head kh * G + r % G), f32 -> fp16 hi / lo, swizzled K-major; each thread of the pair
    //      converts half of the row's 16-byte chunks
    const int r = tid & 127, tq = t0 + r / G, g = r % G;
    const bool row_ok = tq < p.T;
    {
        const float* qrow = p.qkv + (size_t)min(tq, p.T - 1) * p.ld + (size_t)(kh * G + g) * HD;
#pragma unroll
        for (int cc = 0; cc < HD / 16; cc++) {
            const int c = half * (HD / 16) + cc;    // chunk of 8 elements
            const float4 a = *reinterpret_cast<const float4*>(qrow + 8 * c), b = *reinterpret_cast<const float4*>(qrow + 8 * c + 4);
            const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t wh[4], wl[4];
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const float x0 = fminf(fmaxf(f[2 * h], -65504.0f), 65504.0f), x1 = fminf(fmaxf(f[2 * h + 1], -65504.0f), 65504.0f);
                const __half2 hi = __floats2half2_rn(x0, x1);
                const float2 hf = __half22float2(hi);
                wh[h] = *reinterpret_cast<const uint32_t*>(&hi);
                wl[h] = umma_pack_h2(x0 - hf.x, x1 - hf.y);
            }
            const uint32_t off = (uint32_t)(c >> 3) * (128 * 128) + umma_sw128(r, c & 7);
            *reinterpret_cast<uint4*>(sQ + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            *reinterpret_cast<uint4*>(sQ + kQBytes + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem, tS = tmem, tO = tmem + 128u;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;         // TMEM lanes of this warp = query rows
    const int pos_q = p.pos0 + min(tq, p.T - 1);                          // this row attends keys 0 .. pos_q
    const int kv_hi = p.pos0 + min(t0 + TQ, p.T);                         // keys needed by the tile: 0 .. kv_hi - 1
    const int nt = (kv_hi + kAttnTK - 1) / kAttnTK;
    const uint32_t idS = umma_idesc(128, kAttnTK), idO = umma_idesc(128, HD);
    const uint64_t dQh = umma_desc(umma_smem_u32(sQ)), dQl = umma_desc(umma_smem_u32(sQ + kQBytes));
    const uint64_t dKh = umma_desc(umma_smem_u32(sK)), dKl = umma_desc(umma_smem_u32(sK + kKBytes));
    const uint64_t dVh = umma_desc(umma_smem_u32(sV)), dVl = umma_desc(umma_smem_u32(sV + kVBytes));
    const uint64_t dPh = umma_desc(umma_smem_u32(sP)), dPl = umma_desc(umma_smem_u32(sP + kPBytes));
    bool alive = true;
    long long waited = 0;
    uint32_t phK = 0u, phV = 0u, phS = 0u, phPV = 0u;

    auto load_k = [&](int j) {   // thread 0: hi and lo K tiles of key tile j (the previous tile's S MMAs have completed)
        u2_expect_tx(bK, (uint32_t)(2 * kKBytes));
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int s = 0; s < NS; s++)
                u2_tma_2d(umma_smem_u32(sK + h * kKBytes + s * (kAttnTK * 128)), &kmap, s * 64, (h * p.n_kv + kh) * p.P + j * kAttnTK, bK);
    };
    auto load_v = [&](int j) {   // thread 0: hi and lo V^T tiles of key tile j (the previous tile's PV MMAs have completed)
        u2_expect_tx(bV, (uint32_t)(2 * kVBytes));
#pragma unroll
        for (int h = 0; h < 2; h++) u2_tma_2d(umma_smem_u32(sV + h * kVBytes), &vmap, j * kAttnTK, (h * p.n_kv + kh) * HD, bV);
    };
    auto mma_s = [&]() {         // thread 0: S = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const uint64_t da = t == 1 ? dQl : dQh, db = t == 2 ? dKl : dKh;
#pragma unroll
            for (int s = 0; s < NS; s++)
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                    umma_f16(tS, da + (uint64_t)(s * (128 * 128 / 16) + 2 * kk), db + (uint64_t)(s * (kAttnTK * 128 / 16) + 2 * kk), idS,
                             (t > 0 || s > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(bS);
    };
    // this thread's 32 columns of its row of S
    auto ld_s32 = [&](uint32_t (&v)[32]) {
        const uint32_t taddr = tS + lane_base + (uint32_t)(half * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    };

    // ---- pass 1: row maximum and sum of exp (softmax_inplace of ops.rs:1500-1527: max, exp(s - max), sum), per column half
    float m = -INFINITY, l = 0.0f;
    if (tid == 0) load_k(0);
    for (int j = 0; j < nt; j++) {
        if (tid == 0) {
            u2_wait(bK, phK, alive, p.err, 21, waited);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            mma_s();
        }
        phK ^= 1u;
        u2_wait(bS, phS, alive, p.err, 22, waited);
        phS ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && j + 1 < nt) load_k(j + 1);   // lands while the rows are reduced
        {
            uint32_t v[32];
            ld_s32(v);
            float mt = m;
            const int key0 = j * kAttnTK + half * 32;
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const float s = (key0 + q <= pos_q) ? __uint_as_float(v[q]) * p.scale : -INFINITY;
                v[q] = __float_as_uint(s);
                mt = fmaxf(mt, s);
            }
            if (mt != -INFINITY) {
                float add = 0.0f;
#pragma unroll
                for (int q = 0; q < 32; q++) {
                    const float s = __uint_as_float(v[q]);
                    add += (s == -INFINITY) ? 0.0f : __expf(s - mt);
                }
                l = (m == -INFINITY ? 0.0f : l * expf(m - mt)) + add;
                m = mt;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                       // every row of S has been read: the next tile's MMAs may overwrite it
    }
    // the two column halves of a row agree on max and sum
    s_ml[half][r][0] = m;
    s_ml[half][r][1] = l;
    __syncthreads();
    {
        const float m0 = s_ml[0][r][0], l0 = s_ml[0][r][1], m1 = s_ml[1][r][0], l1 = s_ml[1][r][1];
        m = fmaxf(m0, m1);
        l = (m0 == -INFINITY ? 0.0f : l0 * expf(m0 - m)) + (m1 == -INFINITY ? 0.0f : l1 * expf(m1 - m));
    }
    // ---- pass 2: O = sum_p exp(s[p] - max) V[p], normalised at the end
    if (tid == 0) { load_k(0); load_v(0); }
    for (int j = 0; j < nt; j++) {
        if (tid == 0) {
            u2_wait(bK, phK, alive, p.err, 23, waited);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            mma_s();
        }
        phK ^= 1u;
        u2_wait(bS, phS, alive, p.err, 24, waited);
        phS ^= 1u;
        if (j > 0) {                           // the PV MMAs of the previous tile have read P and V
            u2_wait(bPV, phPV, alive, p.err, 25, waited);
            phPV ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {                        // both land while the rows are exponentiated
            if (j + 1 < nt) load_k(j + 1);
            if (j > 0) load_v(j);
        }
        {
            uint32_t v[32];
            ld_s32(v);
            const int key0 = j * kAttnTK + half * 32;
#pragma unroll
            for (int c = 0; c < 4; c++) {      // 8 keys -> one 16-byte chunk of the row, hi and lo
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int q = 8 * c + 2 * h, key = key0 + q;
                    const float e0 = (key <= pos_q) ? __expf(__uint_as_float(v[q]) * p.scale - m) : 0.0f;
                    const float e1 = (key + 1 <= pos_q) ? __expf(__uint_as_float(v[q + 1]) * p.scale - m) : 0.0f;
                    const __half2 hi = __floats2half2_rn(e0, e1);
                    const float2 hf = __half22float2(hi);
                    wh[h] = *reinterpret_cast<const uint32_t*>(&hi);
                    wl[h] = umma_pack_h2(e0 - hf.x, e1 - hf.y);
                }
                const uint32_t off = umma_sw128(r, half * 4 + c);   // chunk 0..7 of the 64-key row
                *reinterpret_cast<uint4*>(sP + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(sP + kPBytes + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            u2_wait(bV, phV, alive, p.err, 26, waited);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // O += P_hi V_hi + P_lo V_hi + P_hi V_lo
#pragma unroll
            for (int t = 0; t < 3; t++) {
                const uint64_t da = t == 1 ? dPl : dPh, db = t == 2 ? dVl : dVh;
#pragma unroll
                for (int kk = 0; kk < kAttnTK / 16; kk++)
                    umma_f16(tO, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idO, (j > 0 || t > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(bPV);
        }
        phV ^= 1u;
    }
    u2_wait(bPV, phPV, alive, p.err, 27, waited);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: O row / sum -> fp16; each thread of the pair stores half of the head dimension
    {
        const float inv = 1.0f / l;
        __half* orow = p.out + (size_t)min(tq, p.T - 1) * p.ldo + (size_t)(kh * G + g) * HD + half * (HD / 2);
#pragma unroll
        for (int c0 = 0; c0 < HD / 2; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tO + lane_base + (uint32_t)(half * (HD / 2) + c0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t w[4];
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const __half2 hh = __halves2half2(f2h_sat(__uint_as_float(v[8 * c + 2 * h]) * inv), f2h_sat(__uint_as_float(v[8 * c + 2 * h + 1]) * inv));
                        w[h] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(orow + c0 + 8 * c) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols) : "memory");
    }
}

// fp16 copies' tensor maps: K16 as {hd, 2 * n_kv * P} (box 64 x 64 keys), Vt16 as {P, 2 * n_kv * hd} (box 64 keys x hd rows), SWIZZLE_128B
inline bool attn_umma_encode(Umma2EncodeFn encode, const __half* k16, const __half* vt16, int n_kv, int hd, int P, CUtensorMap* kmap, CUtensorMap* vmap) {
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t dims[2] = {(cuuint64_t)hd, (cuuint64_t)2 * n_kv * P};   // hi rows, then lo rows
        const cuuint64_t strides[1] = {(cuuint64_t)hd * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)kAttnTK};
        if (encode(kmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)k16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)2 * n_kv * hd};   // hi rows, then lo rows
    const cuuint64_t strides[1] = {(cuuint64_t)P * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)hd};
    return encode(vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)vt16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline size_t attn_umma_smem(int hd) {   // Q, K, V^T, P tiles, each hi + lo
    return (size_t)2 * (128 * hd * 2 + kAttnTK * hd * 2 + hd * kAttnTK * 2 + 128 * kAttnTK * 2) + 1024;
}
inline bool attn_umma_ok(int hd, int n_heads, int n_kv) {
    const int G = n_kv > 0 ? n_heads / n_kv : 0;
    return (hd == 64 || hd == 128) && n_kv > 0 && n_heads % n_kv == 0 && (G == 1 || G == 2 || G == 4 || G == 8);
}

template <int HD>
inline cudaError_t attn_umma_launch_hd(const CUtensorMap& kmap, const CUtensorMap& vmap, const AttnUmmaParams& p, cudaStream_t st) {
    static bool once[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !once[dev]) {
        cudaError_t e = cudaFuncSetAttribute(prefill_attn_umma_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_umma_smem(HD));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) once[dev] = true;
    }
    const int TQ = 128 / p.G;
    dim3 grid((p.T + TQ - 1) / TQ, p.n_kv);
    prefill_attn_umma_kernel<HD><<<grid, 256, attn_umma_smem(HD), st>>>(kmap, vmap, p);
    return cudaGetLastError();
}

}  // namespace b200
