// kv_int8.cuh — INT8 KV cache: the reference's QuantizedKVCache in its Int8 format
// (src/model/kv_quantized.rs:11-20, 143-216 write_kv, 230-270 read_k_range, 366-392 quantize_int8 / dequantize_int8) kept in
// HBM as the bytes + one f32 scale per (kv head, position), and a GQA decode attention that reads it.
//
//   write:  per (kv head, position) row of hd values (K after RoPE, V as projected):
//           max_abs = max |x|, scale = max_abs > 1e-10 ? max_abs / 127 : 1, q = clamp(round(x / scale), -128, 127)
//           (round = half away from zero, Rust f32::round)
//   read:   x' = (q as f32) * scale
//   attention_cached (src/backend/cpu/ops.rs:1479-1537) then runs on the x' rows: s[p] = dot(q, K'[p]) * scale_attn, softmax,
//           out = sum s[p] V'[p] — 1 byte per element and 4 per row instead of 4 bytes per element: the 2.1 GB of f32 KV rows a
//           Llama-3-8B token reads at 8K depth become 0.55 GB.
// Layout per layer: K bytes [n_kv][max_seq][hd] int8, K scales [n_kv][max_seq] f32, the same for V.
// Used by the per-op (graph) decode path when the context was created with the int8 KV format; the megakernels, the
// tensor-core prefill and tensor parallelism keep the f32 cache.
#pragma once
#include "common.cuh"

namespace b200 {

struct RopeKvQ8Params {
    float* q;              // [n_heads * hd] rotated in place
    const float* k;        // [n_kv * hd] raw projection (+bias)
    const float* v;        // [n_kv * hd]
    signed char* k8;       // [n_kv][max_seq][hd]
    signed char* v8;
    float* k_scale;        // [n_kv][max_seq]
    float* v_scale;
    const float* freq;     // [hd/2]
    const int* pos;        // device scalar
    int n_heads, n_kv, hd, max_seq, neox;
    float rope_scale;
};

// quantize_int8 of one row held as VEC values per lane (kv_quantized.rs:366-387); returns the scale on every lane
template <int VEC>
__device__ __forceinline__ float kv_q8_row(const float (&x)[VEC], signed char (&q)[VEC]) {
    float mx = 0.0f;
#pragma unroll
    for (int i = 0; i < VEC; i++) mx = fmaxf(mx, fabsf(x[i]));
    mx = warp_max(mx);
    const float scale = mx > 1e-10f ? __fdiv_rn(mx, 127.0f) : 1.0f;
#pragma unroll
    for (int i = 0; i < VEC; i++) q[i] = (signed char)(int)fminf(fmaxf(roundf(__fdiv_rn(x[i], scale)), -128.0f), 127.0f);
    return scale;
}

// One warp per head unit: n_heads query heads (rotated in place), n_kv key heads (rotated, quantised, stored), n_kv value heads
// (quantised, stored).  Rotation arithmetic as rope_kv_kernel (ops.rs:1216-1337).  HD = 32 * VEC.
template <int VEC>
__global__ void rope_kv_q8_kernel(const RopeKvQ8Params p) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int hd = 32 * VEC, half = hd >> 1;
    if (unit >= p.n_heads + 2 * p.n_kv) return;
    const int pos = *p.pos;
    const float position = (float)pos / p.rope_scale;
    if (unit < p.n_heads + p.n_kv) {
        const bool is_q = unit < p.n_heads;
        const float* src = is_q ? p.q + (size_t)unit * hd : p.k + (size_t)(unit - p.n_heads) * hd;
        float r[VEC];
#pragma unroll
        for (int i = 0; i < VEC; i++) {
            const int e = lane * VEC + i;
            // pair (i0, i1) of element e and which of the two it is
            int pi, second;
            if (p.neox) { pi = e < half ? e : e - half; second = e >= half; }
            else { pi = e >> 1; second = e & 1; }
            const int i0 = p.neox ? pi : 2 * pi, i1 = p.neox ? pi + half : 2 * pi + 1;
            const float theta = position * p.freq[pi];
            const float c = cosf(theta), s = sinf(theta);
            const float x0 = src[i0], x1 = src[i1];
            r[i] = second ? __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c)) : __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        }
        __syncwarp();   // every lane has read its operands before anybody overwrites q in place
        if (is_q) {
#pragma unroll
            for (int i = 0; i < VEC; i++) p.q[(size_t)unit * hd + lane * VEC + i] = r[i];
        } else {
            const int kh = unit - p.n_heads;
            signed char qv[VEC];
            const float scale = kv_q8_row<VEC>(r, qv);
            signed char* dst = p.k8 + ((size_t)kh * p.max_seq + pos) * hd + lane * VEC;
#pragma unroll
            for (int i = 0; i < VEC; i++) dst[i] = qv[i];
            if (lane == 0) p.k_scale[(size_t)kh * p.max_seq + pos] = scale;
        }
    } else {
        const int kh = unit - p.n_heads - p.n_kv;
        float r[VEC];
#pragma unroll
        for (int i = 0; i < VEC; i++) r[i] = p.v[(size_t)kh * hd + lane * VEC + i];
        signed char qv[VEC];
        const float scale = kv_q8_row<VEC>(r, qv);
        signed char* dst = p.v8 + ((size_t)kh * p.max_seq + pos) * hd + lane * VEC;
#pragma unroll
        for (int i = 0; i < VEC; i++) dst[i] = qv[i];
        if (lane == 0) p.v_scale[(size_t)kh * p.max_seq + pos] = scale;
    }
}

struct AttnQ8Params {
    const float* q;            // [n_heads][hd], rotated
    const signed char* k8;     // [n_kv][max_seq][hd]
    const signed char* v8;
    const float* k_scale;      // [n_kv][max_seq]
    const float* v_scale;
    float* part;               // [n_kv][n_splits][G][hd + 2]: unnormalised accumulator, running max, running sum
    float* out;                // [n_heads][hd]
    const int* pos;            // device scalar: kv_len = *pos + 1
    int n_kv, G, max_seq, n_splits;
    float scale;
};

// One warp per (kv head, split): the G query heads of the group share every K / V row read; four positions per step;
// block-wise online softmax (as attn_decode_item).  grid (n_splits, n_kv), 32 threads.
template <int VEC, int GMAX>
__global__ void __launch_bounds__(32) attn_q8_split_kernel(const AttnQ8Params p) {
    constexpr int HD = 32 * VEC, UB = 4;
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x, split = blockIdx.x, kh = blockIdx.y, G = p.G;
    const int kv_len = *p.pos + 1;
    const int chunk = (kv_len + p.n_splits - 1) / p.n_splits;
    const int p_begin = split * chunk, p_end = min(kv_len, p_begin + chunk);
    float q[GMAX][VEC], acc[GMAX][VEC], m[GMAX], l[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; g++) {
        m[g] = -INFINITY;
        l[g] = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            acc[g][v] = 0.0f;
            q[g][v] = g < G ? p.q[(size_t)(kh * G + g) * HD + lane * VEC + v] : 0.0f;
        }
    }
    const signed char* kb = p.k8 + (size_t)kh * p.max_seq * HD + lane * VEC;
    const signed char* vb = p.v8 + (size_t)kh * p.max_seq * HD + lane * VEC;
    const float* ksc = p.k_scale + (size_t)kh * p.max_seq;
    const float* vsc = p.v_scale + (size_t)kh * p.max_seq;
    for (int pos = p_begin; pos < p_end; pos += UB) {
        float kr[UB][VEC], vr[UB][VEC];
#pragma unroll
        for (int u = 0; u < UB; u++) {
            const int pc = min(pos + u, p_end - 1);   // clamped: stays in range, masked below
            const float ks = ksc[pc], vs = vsc[pc];
            if constexpr (VEC == 4) {
                const char4 a = *reinterpret_cast<const char4*>(kb + (size_t)pc * HD), b = *reinterpret_cast<const char4*>(vb + (size_t)pc * HD);
                kr[u][0] = __fmul_rn((float)a.x, ks); kr[u][1] = __fmul_rn((float)a.y, ks); kr[u][2] = __fmul_rn((float)a.z, ks); kr[u][3] = __fmul_rn((float)a.w, ks);
                vr[u][0] = __fmul_rn((float)b.x, vs); vr[u][1] = __fmul_rn((float)b.y, vs); vr[u][2] = __fmul_rn((float)b.z, vs); vr[u][3] = __fmul_rn((float)b.w, vs);
            } else {
                const char2 a = *reinterpret_cast<const char2*>(kb + (size_t)pc * HD), b = *reinterpret_cast<const char2*>(vb + (size_t)pc * HD);
                kr[u][0] = __fmul_rn((float)a.x, ks); kr[u][1] = __fmul_rn((float)a.y, ks);
                vr[u][0] = __fmul_rn((float)b.x, vs); vr[u][1] = __fmul_rn((float)b.y, vs);
            }
        }
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
                    s[u][g] = (pos + u < p_end) ? s[u][g] * p.scale : -INFINITY;
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
            float* dst = p.part + (((size_t)kh * p.n_splits + split) * G + g) * (HD + 2);
#pragma unroll
            for (int v = 0; v < VEC; v++) dst[lane * VEC + v] = acc[g][v];
            if (lane == 0) { dst[HD] = m[g]; dst[HD + 1] = l[g]; }
        }
    }
}

// out[head][:] = sum_s acc_s e^(m_s - m) / sum_s l_s e^(m_s - m), splits in order.  grid n_heads, hd threads.
__global__ void attn_q8_merge_kernel(const AttnQ8Params p, int hd) {
    pdl_launch_dependents();
    pdl_wait();
    const int head = blockIdx.x, kh = head / p.G, g = head - kh * p.G, d = threadIdx.x;
    const float* base = p.part + (((size_t)kh * p.n_splits) * p.G + g) * (hd + 2);
    const size_t stride = (size_t)p.G * (hd + 2);
    float m = -INFINITY;
    for (int s = 0; s < p.n_splits; s++) m = fmaxf(m, base[s * stride + hd]);
    float l = 0.0f, a = 0.0f;
    for (int s = 0; s < p.n_splits; s++) {
        const float ms = base[s * stride + hd];
        if (ms == -INFINITY) continue;   // empty split
        const float c = expf(ms - m);
        l += base[s * stride + hd + 1] * c;
        a += base[s * stride + d] * c;
    }
    p.out[(size_t)head * hd + d] = a / l;
}

}  // namespace b200
