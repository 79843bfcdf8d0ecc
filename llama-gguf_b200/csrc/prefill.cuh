// prefill.cuh — the memory-bound kernels around the tcgen05 dequant-GEMM (gemm_umma.cuh) when a prompt is processed T
// tokens at a time: embedding rows, RoPE + KV-cache write, causal GQA attention over the cache, SwiGLU.
// Same arithmetic as the per-token kernels (embed_kernel, rope_kv_kernel, attn_decode_item) and as the reference
// (LlamaModel::forward src/model/llama.rs:275-362; ops::rope cpu/ops.rs:1216-1337; attention_cached :1479-1537;
// silu_mul_inplace cpu/simd.rs:598-649), applied to T rows.
#pragma once
#include "attention.cuh"
#include "misc.cuh"
#include "quant.cuh"

namespace b200 {

// X[t][:] = dequantised row tokens[t] of the embedding table (bit-exact)
__global__ void prefill_embed_kernel(int type, const uint8_t* __restrict__ table, long long row_bytes, int hidden, const int* __restrict__ tokens,
                                     int vocab, float* __restrict__ X) {
    const int t = blockIdx.x;
    const int token = min(max(tokens[t], 0), vocab - 1);
    const int be = type_block_elems(type), bb = type_block_bytes(type);
    const uint8_t* row = table + (long long)token * row_bytes;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        const int blk = i / be;
        X[(size_t)t * hidden + i] = dequant_elem(type, row + (long long)blk * bb, i - blk * be);
    }
}

struct PrefillRopeParams {
    float* qkv;            // [T][ld]: q | k | v raw projections (+bias); q is rotated in place
    int ld;
    float* k_cache;        // [n_kv][max_seq][hd]
    float* v_cache;
    const float* freq;     // [hd/2]
    int pos0, n_heads, n_kv, hd, max_seq, neox;
    float rope_scale;
};
// one CTA per token: Backend::rope for the token's position pos0 + t, k rows -> cache, v rows -> cache
__global__ void prefill_rope_kv_kernel(const PrefillRopeParams p) {
    const int t = blockIdx.x, pos = p.pos0 + t;
    const int half = p.hd >> 1;
    const int n_pairs = (p.n_heads + p.n_kv) * half, n_v = p.n_kv * p.hd;
    const float position = (float)pos / p.rope_scale;
    float* row = p.qkv + (size_t)t * p.ld;
    const float* kraw = row + (size_t)p.n_heads * p.hd;
    const float* vraw = kraw + (size_t)p.n_kv * p.hd;
    for (int i = threadIdx.x; i < n_pairs + n_v; i += blockDim.x) {
        if (i < n_pairs) {
            const int head = i / half, pi = i - head * half;
            const float theta = position * p.freq[pi];
            const float c = cosf(theta), s = sinf(theta);
            const int i0 = p.neox ? pi : 2 * pi, i1 = p.neox ? pi + half : 2 * pi + 1;
            if (head < p.n_heads) {
                float* d = row + (size_t)head * p.hd;
                const float x0 = d[i0], x1 = d[i1];
                d[i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
                d[i1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
            } else {
                const int kh = head - p.n_heads;
                const float* d = kraw + (size_t)kh * p.hd;
                float* o = p.k_cache + ((size_t)kh * p.max_seq + pos) * p.hd;
                const float x0 = d[i0], x1 = d[i1];
                o[i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
                o[i1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
            }
        } else {
            const int j = i - n_pairs;
            const int kh = j / p.hd, d = j - kh * p.hd;
            p.v_cache[((size_t)kh * p.max_seq + pos) * p.hd + d] = vraw[j];
        }
    }
}

struct PrefillAttnParams {
    const float* qkv;      // [T][ld], q rotated
    int ld;
    const float* k_cache;  // [n_kv][max_seq][hd], positions 0 .. pos0 + T - 1 valid
    const float* v_cache;
    float* out;            // [T][ldo]: [n_heads][hd] per token
    int ldo;
    int pos0, T, n_heads, n_kv, max_seq;
    float scale;
};
// one warp per (token, query head): online softmax over cache positions 0 .. pos0 + t (causal), two positions per step
template <int HD>
__global__ void __launch_bounds__(256) prefill_attn_kernel(const PrefillAttnParams p) {
    constexpr int VEC = HD / 32;
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)p.T * p.n_heads) return;
    const int t = (int)(wid / p.n_heads), h = (int)(wid - (long long)t * p.n_heads);
    const int kh = h / (p.n_heads / p.n_kv);
    const int kv_len = p.pos0 + t + 1;
    float q[VEC], acc[VEC];
    const float* qp = p.qkv + (size_t)t * p.ld + (size_t)h * HD + lane * VEC;
#pragma unroll
    for (int v = 0; v < VEC; v++) { q[v] = qp[v]; acc[v] = 0.0f; }
    const float* kb = p.k_cache + (size_t)kh * p.max_seq * HD + lane * VEC;
    const float* vb = p.v_cache + (size_t)kh * p.max_seq * HD + lane * VEC;
    float m = -INFINITY, l = 0.0f;
    for (int pos = 0; pos < kv_len; pos += 2) {
        const bool two = pos + 1 < kv_len;
        float k0[VEC], k1[VEC], v0[VEC], v1[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            k0[v] = kb[(size_t)pos * HD + v];
            v0[v] = vb[(size_t)pos * HD + v];
            k1[v] = two ? kb[(size_t)(pos + 1) * HD + v] : 0.0f;
            v1[v] = two ? vb[(size_t)(pos + 1) * HD + v] : 0.0f;
        }
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++) { s0 = fmaf(q[v], k0[v], s0); s1 = fmaf(q[v], k1[v], s1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        s0 *= p.scale;
        s1 = two ? s1 * p.scale : -INFINITY;
        const float mn = fmaxf(m, fmaxf(s0, s1));
        const float corr = (m == -INFINITY) ? 0.0f : expf(m - mn);
        const float w0 = expf(s0 - mn), w1 = two ? expf(s1 - mn) : 0.0f;
        l = l * corr + w0 + w1;
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = fmaf(w1, v1[v], fmaf(w0, v0[v], acc[v] * corr));
        m = mn;
    }
    float* op = p.out + (size_t)t * p.ldo + (size_t)h * HD + lane * VEC;
    const float inv = 1.0f / l;
#pragma unroll
    for (int v = 0; v < VEC; v++) op[v] = acc[v] * inv;
}

// g[i] = silu(g[i]) * u[i]  (silu rounded to f32 first, then the product: simd.rs:598-649)
__global__ void prefill_swiglu_kernel(float* __restrict__ g, const float* __restrict__ u, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = g[i];
        g[i] = (x / (1.0f + expf(-x))) * u[i];
    }
}

// the slot's position after T more tokens
__global__ void prefill_advance_kernel(SeqState* st, int T) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        st->pos_cur = st->pos_next + T - 1;
        st->pos_next = st->pos_next + T;
    }
}

}  // namespace b200
