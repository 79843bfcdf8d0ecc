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
    pdl_launch_dependents();
    pdl_wait();
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
    // batched decode: row t is the next token of ITS OWN sequence slot -- position row_pos[t], caches at
    // row_kv[t] + k_off / v_off (floats); nullptr = rows are consecutive positions of one sequence (prefill)
    const int* row_pos;
    float* const* row_kv;
    long long k_off, v_off;
};
// one CTA per token: Backend::rope for the token's position pos0 + t, k rows -> cache, v rows -> cache
__global__ void prefill_rope_kv_kernel(const PrefillRopeParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x, pos = p.row_pos ? p.row_pos[t] : p.pos0 + t;
    float* const k_cache = p.row_pos ? p.row_kv[t] + p.k_off : p.k_cache;
    float* const v_cache = p.row_pos ? p.row_kv[t] + p.v_off : p.v_cache;
    const int half = p.hd >> 1;
    const int n_pairs = (p.n_heads + p.n_kv) * half, n_v = p.n_kv * p.hd;
    const float position = (float)pos / p.rope_scale;
    float* row = p.qkv + (size_t)t * p.ld;
    const float* kraw = row + (size_t)p.n_heads * p.hd;
    const float* vraw = kraw + (size_t)p.n_kv * p.hd;
    // gridDim.y CTAs share a token's items (each item is independent: same arithmetic whoever computes it)
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n_pairs + n_v; i += blockDim.x * gridDim.y) {
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
                float* o = k_cache + ((size_t)kh * p.max_seq + pos) * p.hd;
                const float x0 = d[i0], x1 = d[i1];
                o[i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
                o[i1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
            }
        } else {
            const int j = i - n_pairs;
            const int kh = j / p.hd, d = j - kh * p.hd;
            v_cache[((size_t)kh * p.max_seq + pos) * p.hd + d] = vraw[j];
        }
    }
}

struct PrefillAttnParams {
    const float* qkv;      // [T][ld], q rotated
    int ld;
    const float* k_cache;  // [n_kv][max_seq][hd], positions 0 .. pos0 + T - 1 valid
    const float* v_cache;
    __half* out;           // [T][ldo]: [n_heads][hd] per token, fp16 (input of the O-projection GEMM)
    int ldo;
    int pos0, T, n_heads, n_kv, max_seq;
    float scale;
    const int* row_pos;        // batched decode: see PrefillRopeParams
    float* const* row_kv;
    long long k_off, v_off;
};
// One warp per (token, kv head): the G query heads of the group share every K / V row read (GQA), four cache positions per
// step (16 independent dot products and reductions in flight), block-wise online softmax as in attn_decode_item.
// Causal: token t attends cache positions 0 .. pos0 + t.
template <int HD, int GMAX>
__global__ void __launch_bounds__(128) prefill_attn_kernel(const PrefillAttnParams p) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int VEC = HD / 32, UB = 4;
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)p.T * p.n_kv) return;
    const int t = (int)(wid / p.n_kv), kh = (int)(wid - (long long)t * p.n_kv);
    const int G = p.n_heads / p.n_kv;
    const int kv_len = (p.row_pos ? p.row_pos[t] : p.pos0 + t) + 1;
    float q[GMAX][VEC], acc[GMAX][VEC], m[GMAX], l[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; g++) {
        m[g] = -INFINITY;
        l[g] = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            acc[g][v] = 0.0f;
            q[g][v] = g < G ? p.qkv[(size_t)t * p.ld + (size_t)(kh * G + g) * HD + lane * VEC + v] : 0.0f;
        }
    }
    const float* kb = (p.row_pos ? p.row_kv[t] + p.k_off : p.k_cache) + (size_t)kh * p.max_seq * HD + lane * VEC;
    const float* vb = (p.row_pos ? p.row_kv[t] + p.v_off : p.v_cache) + (size_t)kh * p.max_seq * HD + lane * VEC;
    // the K / V rows of the next four positions are in flight while these four are processed (two register buffers)
    auto load = [&](int pos, float (&kk)[UB][VEC], float (&vv)[UB][VEC]) {
#pragma unroll
        for (int u = 0; u < UB; u++) {
            const int pc = min(pos + u, kv_len - 1);   // clamped: stays in range, masked below
            if constexpr (VEC == 4) {
                const float4 a = *reinterpret_cast<const float4*>(kb + (size_t)pc * HD), b = *reinterpret_cast<const float4*>(vb + (size_t)pc * HD);
                kk[u][0] = a.x; kk[u][1] = a.y; kk[u][2] = a.z; kk[u][3] = a.w;
                vv[u][0] = b.x; vv[u][1] = b.y; vv[u][2] = b.z; vv[u][3] = b.w;
            } else {
                const float2 a = *reinterpret_cast<const float2*>(kb + (size_t)pc * HD), b = *reinterpret_cast<const float2*>(vb + (size_t)pc * HD);
                kk[u][0] = a.x; kk[u][1] = a.y;
                vv[u][0] = b.x; vv[u][1] = b.y;
            }
        }
    };
    float kn[UB][VEC], vn[UB][VEC];
    load(0, kn, vn);
    for (int pos = 0; pos < kv_len; pos += UB) {
        float kr[UB][VEC], vr[UB][VEC];
#pragma unroll
        for (int u = 0; u < UB; u++)
#pragma unroll
            for (int v = 0; v < VEC; v++) { kr[u][v] = kn[u][v]; vr[u][v] = vn[u][v]; }
        if (pos + UB < kv_len) load(pos + UB, kn, vn);
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < UB; u++)
#pragma unroll
                for (int g = 0; g < GMAX; g++) s[u][g] += __shfl_xor_sync(0xffffffffu, s[u][g], o);
#pragma unroll
        for (int g = 0; g < GMAX; g++) {
            if (g < G) {
                float mb = -INFINITY;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    s[u][g] = (pos + u < kv_len) ? s[u][g] * p.scale : -INFINITY;
                    mb = fmaxf(mb, s[u][g]);
                }
                const float mn = fmaxf(m[g], mb);
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
    }
#pragma unroll
    for (int g = 0; g < GMAX; g++) {
        if (g < G) {
            const float inv = 1.0f / l[g];
#pragma unroll
            for (int v = 0; v < VEC; v++) p.out[(size_t)t * p.ldo + (size_t)(kh * G + g) * HD + lane * VEC + v] = f2h_sat(acc[g][v] * inv);
        }
    }
}

// RMSNorm of T rows, output rounded to fp16 for the tensor-core GEMM that consumes it (same f32 arithmetic as
// rms_norm_rows_kernel: ss = sum x^2, inv = 1/sqrt(ss/n + eps), (x*inv)*w; simd.rs:847-899)
// Rows of up to 256 * kRmsMaxV elements are held in registers: all loads of a thread are issued before the first is used (a row is
// 16 dependent-latency trips otherwise: 15 us per launch at 32 rows), the sums run in the SAME order as before (bit-identical).
constexpr int kRmsMaxV = 32;
__global__ void __launch_bounds__(256) prefill_rms_norm_kernel(const float* x, const float* w, float eps, __half* out, int n) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[8];
    const float* xr = x + (size_t)blockIdx.x * n;
    __half* orow = out + (size_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float ss = 0.0f;
    if (n <= 256 * kRmsMaxV) {
        float v[kRmsMaxV], wv[kRmsMaxV];
#pragma unroll
        for (int k = 0; k < kRmsMaxV; k++) {
            const int i = threadIdx.x + 256 * k;
            v[k] = i < n ? xr[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < kRmsMaxV; k++) {
            const int i = threadIdx.x + 256 * k;
            wv[k] = i < n ? w[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < kRmsMaxV; k++)
            if (threadIdx.x + 256 * k < n) ss = fmaf(v[k], v[k], ss);
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.0f;
        for (int k = 0; k < 8; k++) tot += red[k];
        const float inv = 1.0f / sqrtf(tot / (float)n + eps);
#pragma unroll
        for (int k = 0; k < kRmsMaxV; k++) {
            const int i = threadIdx.x + 256 * k;
            if (i < n) orow[i] = f2h_sat(__fmul_rn(__fmul_rn(v[k], inv), wv[k]));
        }
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) ss = fmaf(xr[i], xr[i], ss);
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.0f;
    for (int k = 0; k < 8; k++) tot += red[k];
    const float inv = 1.0f / sqrtf(tot / (float)n + eps);
    for (int i = threadIdx.x; i < n; i += blockDim.x) orow[i] = f2h_sat(__fmul_rn(__fmul_rn(xr[i], inv), w[i]));
}

// g[i] = silu(g[i]) * u[i]  (silu rounded to f32 first, then the product: simd.rs:598-649)
__global__ void prefill_swiglu_kernel(const float* __restrict__ g, const float* __restrict__ u, __half* __restrict__ out, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = g[i];
        out[i] = f2h_sat((x / (1.0f + expf(-x))) * u[i]);
    }
}

// the slot's position after T more tokens
__global__ void prefill_advance_kernel(SeqState* st, int T) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        st->pos_cur = st->pos_next + T - 1;
        st->pos_next = st->pos_next + T;
    }
}

// batched decode: every row's slot advances by one token
__global__ void prefill_advance_rows_kernel(SeqState* const* st, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        SeqState* s = st[i];
        s->pos_cur = s->pos_next;
        s->pos_next = s->pos_next + 1;
    }
}

}  // namespace b200
