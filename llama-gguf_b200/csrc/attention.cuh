// attention.cuh — RoPE + KV-cache write, and GQA split-KV decode attention.
//
// Replaces rope_single_pos / update_kv_cache / flash_attention_cached of the existing
// CUDA backend (src/backend/cuda/kernels.rs:379-441, 800-824, 1395-1460) and follows the
// CPU semantics of Backend::rope (src/backend/cpu/ops.rs:1216-1337), the cache write
// (src/model/layers.rs:580-600) and Backend::attention_cached (cpu/ops.rs:1479-1537).
//
// KV cache layout per layer (src/model/mod.rs:83-108): [n_kv_heads][max_seq][head_dim] f32.
//
// attn_decode: one CTA per (kv head, KV split).  The G = n_heads/n_kv_heads query heads
// that share a kv head are processed together so every K/V row is read from HBM once
// per group, not once per query head.  A warp owns a strided subset of the split's
// positions; each lane holds head_dim/32 contiguous floats of the row (coalesced 128-bit
// loads), scores are warp-shuffle reductions, softmax is online (running max / sum) in
// f32.  Warps combine through shared memory, splits through a global scratch; the last
// CTA to arrive for a kv head (atomic ticket) does the final merge in a fixed order, so
// results are run-to-run deterministic.
//
// Deviation (documented in DESIGN.md): the reference skips value rows whose probability
// is <= 1e-8 (cpu/ops.rs:1529); this kernel includes them.  The dropped mass is at most
// kv_len * 1e-8 of the output — far below the 1e-3 parity tolerance.
#pragma once
#include "common.cuh"

namespace b200 {

struct RopeKvParams {
    float* q;              // [n_heads * hd] in place
    const float* k;        // [n_kv * hd] raw projection (+bias)
    const float* v;        // [n_kv * hd]
    float* k_cache;        // [n_kv][max_seq][hd]
    float* v_cache;
    const float* freq;     // [hd/2] = 1 / base^(2i/hd), computed on the host with libm powf
    const int* pos;        // device scalar: position of the current token
    int n_heads, n_kv, hd, max_seq, neox;
    float rope_scale;
    long long kv_head_stride, kv_pos_stride;   // floats; 0 = the reference layout [n_kv][max_seq][hd]
};
// offset (floats) of row (kv head, position) in a K or V cache
template <class P>
__device__ __forceinline__ size_t kv_row(const P& p, int kh, int pos, int hd) {
    return p.kv_pos_stride ? (size_t)kh * (size_t)p.kv_head_stride + (size_t)pos * (size_t)p.kv_pos_stride
                           : ((size_t)kh * p.max_seq + pos) * hd;
}

// Backend::rope + cache write.  One thread per rotated pair, then one per v element.
__global__ void rope_kv_kernel(const RopeKvParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int pos = *p.pos;
    const int half = p.hd >> 1;
    const int n_pairs = (p.n_heads + p.n_kv) * half;
    const int n_v = p.n_kv * p.hd;
    const float position = (float)pos / p.rope_scale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs + n_v; i += gridDim.x * blockDim.x) {
        if (i < n_pairs) {
            const int head = i / half, pi = i - head * half;
            const float theta = position * p.freq[pi];
            const float c = cosf(theta), s = sinf(theta);
            const int i0 = p.neox ? pi : 2 * pi;
            const int i1 = p.neox ? pi + half : 2 * pi + 1;
            if (head < p.n_heads) {
                float* d = p.q + (size_t)head * p.hd;
                const float x0 = d[i0], x1 = d[i1];
                d[i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
                d[i1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
            } else {
                const int kh = head - p.n_heads;
                const float* d = p.k + (size_t)kh * p.hd;
                float* o = p.k_cache + kv_row(p, kh, pos, p.hd);
                const float x0 = d[i0], x1 = d[i1];
                o[i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
                o[i1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
            }
        } else {
            const int j = i - n_pairs;
            const int kh = j / p.hd, d = j - kh * p.hd;
            p.v_cache[kv_row(p, kh, pos, p.hd) + d] = p.v[j];
        }
    }
}

constexpr int kAttnWarps = 8;
constexpr int kAttnThreads = kAttnWarps * kWarp;

struct AttnParams {
    const float* q;        // [n_heads][hd], RoPE applied
    const float* k_cache;  // [n_kv][max_seq][hd]
    const float* v_cache;
    float* out;            // [n_heads][hd]
    float* part;           // scratch [n_kv][n_splits][G][hd + 2]
    unsigned int* tickets; // [n_kv], zero on entry, left zero on exit
    const int* pos;        // device scalar (kv_len = *pos + 1) or nullptr
    int kv_len_fixed;      // used when pos == nullptr
    int n_kv, G, max_seq, n_splits;
    int min_chunk;         // KV positions per split below which a context is not split further (0: kAttnMinChunk)
    float scale;
    // optional fused RoPE + KV-cache write (per-token megakernel): q | k | v raw from the QKV GEMV (+bias), rotated
    // here exactly as rope_kv_kernel does; the split that owns position `pos` also writes the cache rows
    const float* qkv_raw;  // nullptr: q is already rotated, the cache already holds position pos
    const float* freq;     // [hd/2]
    float rope_scale;
    int neox, n_heads;
    // optional: also write the staged (int8 planes) form of the output vector for the GEMV that consumes it
    // (gemv_mma.cuh: stage_out32); stage_K = n_heads * hd
    uint8_t* stage_out;
    int stage_K;
    // optional strides (floats) of a position-major cache ([pos][k|v][n_kv][hd]); 0 = the reference layout
    // [n_kv][max_seq][hd] (model/mod.rs:83-108), which is what the engine uses (position-major measured no faster)
    long long kv_head_stride, kv_pos_stride;
    unsigned long long* dbg;   // optional [grid][8] globaltimer stamps of thread 0 (debug timelines)
};
__device__ __forceinline__ void attn_stamp(const AttnParams& p, int i) {
    if (p.dbg && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.dbg[(size_t)blockIdx.x * 8 + i] = t;
    }
}

// KV positions per split: short contexts use few CTAs (the grid is sized for max_seq; surplus CTAs exit at once)
#ifndef B200_ATTN_MIN_CHUNK
#define B200_ATTN_MIN_CHUNK 32
#endif
constexpr int kAttnMinChunk = B200_ATTN_MIN_CHUNK;
__device__ __forceinline__ int attn_eff_splits(int kv_len, int n_splits, int min_chunk) {
    const int mc = min_chunk > 0 ? min_chunk : kAttnMinChunk;
    return max(1, min(n_splits, (kv_len + mc - 1) / mc));
}

// defined in gemv_mma.cuh (stage_out32 without a weight); a warp holds 32 consecutive elements of the output vector
__device__ __forceinline__ void attn_stage_out(float val, int j, int K, uint8_t* xg);

// merge (m, l, acc) <- (m, l, acc) (+) (m2, l2, acc2)
__device__ __forceinline__ void softmax_merge_scale(float m, float m2, float& ca, float& cb, float& mo) {
    mo = fmaxf(m, m2);
    ca = (m == -INFINITY) ? 0.0f : expf(m - mo);
    cb = (m2 == -INFINITY) ? 0.0f : expf(m2 - mo);
}

// Barrier of the NT threads that execute an item: named barrier 1, so that the streamed megakernel (stream.cuh) can
// run an extra producer warp that never joins (in the stand-alone kernels NT is the whole CTA).
template <int NT>
__device__ __forceinline__ void attn_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// One (kv head, KV split) work item, executed by the NW warps of a CTA.  sm: (2*NW*GMAX + NW*GMAX*HD + GMAX*HD)
// floats (>= 64*GMAX + GMAX); s_ticket: one shared word.
template <int HD, int GMAX, int NW>
__device__ __forceinline__ void attn_decode_item(const AttnParams& p, int kh, int split, int kv_len, float* sm,
                                                 unsigned int* s_ticket, const float* s_rope = nullptr) {
    constexpr int VEC = HD / 32;  // floats per lane per row
    constexpr int NT = NW * 32;
    float* s_m = sm;                      // [warps][GMAX]
    float* s_l = sm + NW * GMAX;          // [warps][GMAX]
    float* s_acc = sm + 2 * NW * GMAX;    // [warps][GMAX][HD]
    float* s_q = s_acc + NW * GMAX * HD;  // [GMAX][HD] (fused RoPE only)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = p.G;
    const int ns = attn_eff_splits(kv_len, p.n_splits, p.min_chunk);
    if (split >= ns) return;
    int chunk = (kv_len + ns - 1) / ns;
    chunk = (chunk + NW - 1) / NW * NW;
    const int start = split * chunk;
    const int end = min(kv_len, start + chunk);

    if (p.qkv_raw) {  // Backend::rope (cpu/ops.rs:1216-1337) + cache write (layers.rs:580-600)
        // s_rope (per-token table of the megakernel: cos[pi], sin[pi] at [pi], [HD/2 + pi]) holds exactly the values
        // computed below; every global load of this block is issued before the first use (one round trip)
        const int pos = kv_len - 1, half = HD / 2;
        const float position = (float)pos / p.rope_scale;
        const bool own = pos >= start && pos < end;   // this split writes the cache rows of the new position
        const float* kraw = p.qkv_raw + (size_t)p.n_heads * HD + (size_t)kh * HD;
        const float* vraw = p.qkv_raw + (size_t)(p.n_heads + p.n_kv) * HD + (size_t)kh * HD;
        for (int idx = threadIdx.x; idx < G * half; idx += NT) {
            const int g = idx / half, pi = idx - g * half;
            const int i0 = p.neox ? pi : 2 * pi, i1 = p.neox ? pi + half : 2 * pi + 1;
            const float* d = p.qkv_raw + (size_t)(kh * G + g) * HD;
            const float x0 = d[i0], x1 = d[i1];
            const bool kk = own && idx < half;        // the first `half` threads also rotate the k row
            const float k0 = kk ? kraw[i0] : 0.f, k1 = kk ? kraw[i1] : 0.f;
            const bool vv = own && idx < HD;
            const float v0 = vv ? vraw[idx] : 0.f;
            float c, sn;
            if (s_rope) { c = s_rope[pi]; sn = s_rope[half + pi]; }
            else { const float theta = position * p.freq[pi]; c = cosf(theta); sn = sinf(theta); }
            s_q[g * HD + i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, sn));
            s_q[g * HD + i1] = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
            if (kk) {
                float* ko = const_cast<float*>(p.k_cache) + kv_row(p, kh, pos, HD);
                ko[i0] = __fsub_rn(__fmul_rn(k0, c), __fmul_rn(k1, sn));
                ko[i1] = __fadd_rn(__fmul_rn(k0, sn), __fmul_rn(k1, c));
            }
            if (vv) const_cast<float*>(p.v_cache)[kv_row(p, kh, pos, HD) + idx] = v0;
        }
        if (own && G * half < HD) {   // fewer rotated q pairs than v elements (G == 1): the rest of the v row
            for (int d = G * half + threadIdx.x; d < HD; d += NT) const_cast<float*>(p.v_cache)[kv_row(p, kh, pos, HD) + d] = vraw[d];
        }
        attn_sync<NT>();
    }
    attn_stamp(p, 2);

    float q[GMAX][VEC], acc[GMAX][VEC], m[GMAX], l[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; g++) {
        m[g] = -INFINITY;
        l[g] = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            acc[g][v] = 0.0f;
            q[g][v] = (g < G) ? (p.qkv_raw ? s_q[g * HD + lane * VEC + v] : p.q[((size_t)(kh * G + g)) * HD + lane * VEC + v]) : 0.0f;
        }
    }
    const float* kb = p.k_cache + kv_row(p, kh, 0, HD) + lane * VEC;
    const float* vb = p.v_cache + kv_row(p, kh, 0, HD) + lane * VEC;
    const size_t pstride = p.kv_pos_stride ? (size_t)p.kv_pos_stride : (size_t)HD;

    // A warp owns positions start + warp + i * NW.  Batches of UB positions: all dot products and their warp reductions
    // of a batch are independent, one running-max update per batch (block-wise online softmax), and the K/V rows of
    // the next batch are in flight while this one is computed (two register buffers).
    constexpr int UB = 4;
    constexpr int STEP = NW * UB;
    auto load = [&](int pos0, float (&kr)[UB][VEC], float (&vr)[UB][VEC]) {
#pragma unroll
        for (int u = 0; u < UB; u++) {
            const int pp = pos0 + u * NW;
            const int pc = pp < end ? pp : pos0;  // clamp: loads stay in range, result discarded
            if constexpr (VEC == 4) {
                const float4 a = *reinterpret_cast<const float4*>(kb + (size_t)pc * pstride);
                const float4 b = *reinterpret_cast<const float4*>(vb + (size_t)pc * pstride);
                kr[u][0] = a.x; kr[u][1] = a.y; kr[u][2] = a.z; kr[u][3] = a.w;
                vr[u][0] = b.x; vr[u][1] = b.y; vr[u][2] = b.z; vr[u][3] = b.w;
            } else {
                const float2 a = *reinterpret_cast<const float2*>(kb + (size_t)pc * pstride);
                const float2 b = *reinterpret_cast<const float2*>(vb + (size_t)pc * pstride);
                kr[u][0] = a.x; kr[u][1] = a.y;
                vr[u][0] = b.x; vr[u][1] = b.y;
            }
        }
    };
#ifdef B200_ATTN_PROBE
    long long pr_t0 = clock64();
    int pr_n = 0;
#define ATTN_PROBE(i) do { if (p.dbg && threadIdx.x == 0 && pr_n == 1) p.dbg[(size_t)blockIdx.x * 8 + (i)] = (unsigned long long)(clock64() - pr_t0); } while (0)
#else
#define ATTN_PROBE(i) do {} while (0)
#endif
    auto compute = [&](int pos0, const float (&kr)[UB][VEC], const float (&vr)[UB][VEC]) {
#ifdef B200_ATTN_PROBE
        pr_n++;
        if (pr_n == 1) pr_t0 = clock64();
#endif
        float s[UB][GMAX];
#pragma unroll
        for (int u = 0; u < UB; u++)
#pragma unroll
            for (int g = 0; g < GMAX; g++) {
                float d = 0.0f;
#pragma unroll
                for (int v = 0; v < VEC; v++) d = fmaf(q[g][v], kr[u][v], d);
                s[u][g] = d;
            }
        ATTN_PROBE(4);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < UB; u++)
#pragma unroll
                for (int g = 0; g < GMAX; g++) s[u][g] += __shfl_xor_sync(0xffffffffu, s[u][g], o);
        ATTN_PROBE(5);
#pragma unroll
        for (int g = 0; g < GMAX; g++) {
            if (g < G) {
                float mb = -INFINITY;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    s[u][g] = (pos0 + u * NW < end) ? s[u][g] * p.scale : -INFINITY;   // warp-uniform
                    mb = fmaxf(mb, s[u][g]);
                }
                const float mn = fmaxf(m[g], mb);   // finite: the first position of a batch is always valid
                const float corr = (m[g] == -INFINITY) ? 0.0f : expf(m[g] - mn);
                float w[UB], ws = 0.0f;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    w[u] = (s[u][g] == -INFINITY) ? 0.0f : expf(s[u][g] - mn);
                    ws += w[u];
                }
                l[g] = l[g] * corr + ws;
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    float a = acc[g][v] * corr;
#pragma unroll
                    for (int u = 0; u < UB; u++) a = fmaf(w[u], vr[u][v], a);
                    acc[g][v] = a;
                }
                m[g] = mn;
            }
        }
        ATTN_PROBE(6);
    };
    {
        float kA[UB][VEC], vA[UB][VEC], kB[UB][VEC], vB[UB][VEC];
        int pos0 = start + warp;
        if (pos0 < end) load(pos0, kA, vA);
        while (pos0 < end) {
            if (pos0 + STEP < end) load(pos0 + STEP, kB, vB);
            compute(pos0, kA, vA);
            pos0 += STEP;
            if (pos0 >= end) break;
            if (pos0 + STEP < end) load(pos0 + STEP, kA, vA);
            compute(pos0, kB, vB);
            pos0 += STEP;
        }
    }

    // ---- combine the warps of this CTA ----
    attn_stamp(p, 3);
#pragma unroll
    for (int g = 0; g < GMAX; g++) {
        if (g < G) {
            if (lane == 0) {
                s_m[warp * GMAX + g] = m[g];
                s_l[warp * GMAX + g] = l[g];
            }
#pragma unroll
            for (int v = 0; v < VEC; v++) s_acc[(warp * GMAX + g) * HD + lane * VEC + v] = acc[g][v];
        }
    }
    attn_sync<NT>();
    const int part_stride = HD + 2;
    float* my_part = p.part + ((size_t)(kh * p.n_splits + split) * G) * part_stride;  // slots sized for n_splits
    // merge coefficients once per (warp, head): thread g turns s_m[w][g] into exp(m_w - M) and leaves (M, L) in
    // s_l[0][g], s_l[1][g] (an expf per element and warp in the loop below was ~2 us of dependent math per item)
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; w++) M = fmaxf(M, s_m[w * GMAX + g]);
        float L = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const float mw = s_m[w * GMAX + g];
            const float c = (mw == -INFINITY) ? 0.0f : expf(mw - M);
            L += s_l[w * GMAX + g] * c;
            s_m[w * GMAX + g] = c;
        }
        s_l[0 * GMAX + g] = M;
        s_l[1 * GMAX + g] = L;
    }
    attn_sync<NT>();
    for (int idx = threadIdx.x; idx < G * HD; idx += NT) {
        const int g = idx / HD, d = idx - g * HD;
        const float M = s_l[0 * GMAX + g], L = s_l[1 * GMAX + g];
        float A = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) A += s_acc[(w * GMAX + g) * HD + d] * s_m[w * GMAX + g];
        if (ns == 1) {  // a single split: this is the answer (no scratch, no ticket)
            const float o = A / L;
            p.out[(kh * G + g) * HD + d] = o;
            if (p.stage_out) attn_stage_out(o, (kh * G + g) * HD + d, p.stage_K, p.stage_out);
        } else {
            my_part[g * part_stride + d] = A;
            if (d == 0) {
                my_part[g * part_stride + HD] = M;
                my_part[g * part_stride + HD + 1] = L;
            }
        }
    }
    if (ns == 1) return;

    // ---- last CTA of this kv head merges the splits ----
    attn_stamp(p, 4);
    attn_sync<NT>();
    if (threadIdx.x == 0) {  // release: this CTA's partials are visible; acquire: so are the others'
        unsigned int t;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(p.tickets + kh) : "memory");
        *s_ticket = t;
    }
    attn_sync<NT>();
    attn_stamp(p, 5);
    if (*s_ticket != (unsigned)(ns - 1)) return;
    // all partials of this kv head -> shared memory in one round trip, then merge from there
    const float* parts = p.part + (size_t)kh * p.n_splits * G * part_stride;
    const int n_part = ns * G * part_stride;     // contiguous: splits 0..ns-1 of this head
    float* s_p = sm;                             // [ns][G][HD + 2]   (fits: ns <= NW * GMAX * HD / (G * (HD+2)) is checked by the host)
    // batches of 8 independent loads per thread (a plain copy loop keeps one load in flight per iteration: ~0.7 us each)
    for (int i0 = threadIdx.x; i0 < n_part; i0 += NT * 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (i0 + k * NT < n_part) ? __ldcg(parts + i0 + k * NT) : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (i0 + k * NT < n_part) s_p[i0 + k * NT] = v[k];
    }
    attn_sync<NT>();
    // coefficients once per (split, head): thread g turns the split maxima into exp(m_s - M), leaves L in split 0's slot
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        float M = -INFINITY;
        for (int s = 0; s < ns; s++) M = fmaxf(M, s_p[(s * G + g) * part_stride + HD]);
        float L = 0.0f;
        for (int s = 0; s < ns; s++) {
            float* ps = s_p + (s * G + g) * part_stride;
            const float ms = ps[HD];
            const float c = (ms == -INFINITY) ? 0.0f : expf(ms - M);
            L += ps[HD + 1] * c;
            ps[HD] = c;
        }
        s_p[(0 * G + g) * part_stride + HD + 1] = L;
    }
    attn_sync<NT>();
    for (int idx = threadIdx.x; idx < G * HD; idx += NT) {
        const int g = idx / HD, d = idx - g * HD;
        const float L = s_p[(0 * G + g) * part_stride + HD + 1];
        float A = 0.0f;
        for (int s = 0; s < ns; s++) {
            const float* ps = s_p + (s * G + g) * part_stride;
            A += ps[d] * ps[HD];
        }
        const float o = A / L;
        p.out[(kh * G + g) * HD + d] = o;
        if (p.stage_out) attn_stage_out(o, (kh * G + g) * HD + d, p.stage_K, p.stage_out);
    }
    attn_stamp(p, 6);
    if (threadIdx.x == 0) p.tickets[kh] = 0;  // ready for the next launch / graph replay
}

template <int HD, int GMAX>
__global__ void __launch_bounds__(kAttnThreads) attn_decode_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ unsigned int s_ticket;
    pdl_launch_dependents();
    pdl_wait();
    const int kv_len = p.pos ? (*p.pos + 1) : p.kv_len_fixed;
    attn_decode_item<HD, GMAX, kAttnWarps>(p, blockIdx.x, blockIdx.y, kv_len, sm, &s_ticket);
}

// floats of shared memory attn_decode_item needs
__host__ __device__ inline size_t attn_item_floats(int hd, int gmax, int nw, int n_splits, int G) {
    size_t a = (size_t)(2 * nw * gmax + nw * gmax * hd + gmax * hd), b = (size_t)n_splits * G * (hd + 2);
    return a > b ? a : b;
}
inline size_t attn_smem_bytes(int hd, int gmax, int n_splits, int G) { return attn_item_floats(hd, gmax, kAttnWarps, n_splits, G) * sizeof(float); }

// Backend::attention (cpu/ops.rs:1353-1470): causal, q[n_heads][seq][hd], k/v[n_kv][kv_len][hd].
// Compatibility surface only (the model path uses attention_cached): one warp per (head, query).
__global__ void attention_full_kernel(const float* q, const float* k, const float* v, float* out, int n_heads,
                                      int n_kv, int seq_len, int kv_len, int hd, float scale) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= n_heads * seq_len) return;
    const int head = wid / seq_len, s = wid - head * seq_len;
    const int kvh = head / (n_heads / n_kv);
    const int q_abs = max(kv_len - seq_len, 0) + s;
    const float* qv = q + ((size_t)head * seq_len + s) * hd;
    float m = -INFINITY, l = 0.0f;
    float acc[8];  // hd <= 256
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0f;
    for (int pp = 0; pp < kv_len && pp <= q_abs; pp++) {
        const float* kv = k + ((size_t)kvh * kv_len + pp) * hd;
        const float* vv = v + ((size_t)kvh * kv_len + pp) * hd;
        float d = 0.0f;
        for (int i = lane; i < hd; i += 32) d = fmaf(qv[i], kv[i], d);
        d = warp_sum(d) * scale;
        const float mn = fmaxf(m, d);
        const float corr = (m == -INFINITY) ? 0.0f : expf(m - mn);
        const float w = expf(d - mn);
        l = l * corr + w;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int dd = lane + 32 * i;
            if (dd < hd) acc[i] = fmaf(w, vv[dd], acc[i] * corr);
        }
        m = mn;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int dd = lane + 32 * i;
        if (dd < hd) out[((size_t)head * seq_len + s) * hd + dd] = acc[i] / l;
    }
}

}  // namespace b200
