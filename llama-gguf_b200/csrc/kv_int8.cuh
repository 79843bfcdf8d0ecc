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
// (quantised, stored).  Rotation arithmetic as rope_kv_kernel (ops.rs:1216-1337).  HD = 32 * VEC.  `q`, `k`, `v` are the rows of ONE
// token, `pos` its position.
template <int VEC>
__device__ __forceinline__ void rope_kv_q8_unit(const RopeKvQ8Params& p, float* q, const float* k, const float* v, int pos, int unit, int lane) {
    const int hd = 32 * VEC, half = hd >> 1;
    const float position = (float)pos / p.rope_scale;
    if (unit < p.n_heads + p.n_kv) {
        const bool is_q = unit < p.n_heads;
        const float* src = is_q ? q + (size_t)unit * hd : k + (size_t)(unit - p.n_heads) * hd;
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
            for (int i = 0; i < VEC; i++) q[(size_t)unit * hd + lane * VEC + i] = r[i];
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
        for (int i = 0; i < VEC; i++) r[i] = v[(size_t)kh * hd + lane * VEC + i];
        signed char qv[VEC];
        const float scale = kv_q8_row<VEC>(r, qv);
        signed char* dst = p.v8 + ((size_t)kh * p.max_seq + pos) * hd + lane * VEC;
#pragma unroll
        for (int i = 0; i < VEC; i++) dst[i] = qv[i];
        if (lane == 0) p.v_scale[(size_t)kh * p.max_seq + pos] = scale;
    }
}

template <int VEC>
__global__ void rope_kv_q8_kernel(const RopeKvQ8Params p) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (unit >= p.n_heads + 2 * p.n_kv) return;
    rope_kv_q8_unit<VEC>(p, p.q, p.k, p.v, *p.pos, unit, lane);
}

// The same for the T rows of a prompt chunk (tensor-core prefill): row t = [q | k | v] of the token at position pos0 + t, `ld`
// floats apart (p.q = the first row; p.k / p.v / p.pos unused).  grid (ceil(units / 4), T), 128 threads.
template <int VEC>
__global__ void prefill_rope_kv_q8_kernel(const RopeKvQ8Params p, int ld, int pos0) {
    const int lane = threadIdx.x & 31, unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, t = blockIdx.y;
    if (unit >= p.n_heads + 2 * p.n_kv) return;
    const int hd = 32 * VEC;
    float* q = p.q + (size_t)t * ld;
    const float* k = q + (size_t)p.n_heads * hd;
    const float* v = k + (size_t)p.n_kv * hd;
    rope_kv_q8_unit<VEC>(p, q, k, v, pos0 + t, unit, lane);
}

// fp16 hi / lo pairs of one layer's DEQUANTISED cache rows 0 .. kv_end-1 for the tensor-core prefill attention: the int8 peer of
// prefill_kv16_kernel (attn_umma.cuh), same output layout (K16[kv][pos][hd], Vt16[kv][hd][pos], lo halves after the hi halves).
// The attention of the chunk then sees q * scale for every position, its own included -- what attention over a
// QuantizedKVCache sees (kv_quantized.rs:230-310).  grid (kv_pad / 64, n_kv), 256 threads.
template <int HD>
__global__ void __launch_bounds__(256) prefill_kv16_q8_kernel(const signed char* __restrict__ k8, const signed char* __restrict__ v8,
                                                              const float* __restrict__ k_scale, const float* __restrict__ v_scale, int max_seq,
                                                              int kv_end, int P, __half* __restrict__ k16, __half* __restrict__ vt16) {
    __shared__ __half tile[64][HD + 2], tile_lo[64][HD + 2];
    const int kh = blockIdx.y, n_kv = gridDim.y, p0 = blockIdx.x * 64;
    const signed char* ks = k8 + (size_t)kh * max_seq * HD;
    const signed char* vs = v8 + (size_t)kh * max_seq * HD;
    const float* ksc = k_scale + (size_t)kh * max_seq;
    const float* vsc = v_scale + (size_t)kh * max_seq;
    for (int i = threadIdx.x; i < 64 * HD; i += 256) {
        const int pr = i / HD, d = i - pr * HD, pos = p0 + pr;
        const bool ok = pos < kv_end;
        const float k = ok ? fminf(fmaxf(__fmul_rn((float)ks[(size_t)pos * HD + d], ksc[pos]), -65504.0f), 65504.0f) : 0.0f;
        const float v = ok ? fminf(fmaxf(__fmul_rn((float)vs[(size_t)pos * HD + d], vsc[pos]), -65504.0f), 65504.0f) : 0.0f;
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

// 16 int8 values (one 16-byte load) -> 16 floats without I2F (quarter rate): x ^ 0x80 is the unsigned byte x + 128, PRMT drops it
// into the mantissa of 2^23 (0x4B000000 | u = 2^23 + u exactly), one FADD removes 2^23 + 128.
__device__ __forceinline__ void kv_q8_unpack16(const int4 raw, float (&f)[16]) {
    const uint32_t w[4] = {(uint32_t)raw.x ^ 0x80808080u, (uint32_t)raw.y ^ 0x80808080u, (uint32_t)raw.z ^ 0x80808080u, (uint32_t)raw.w ^ 0x80808080u};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        f[4 * i + 0] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7650)) - 8388736.0f;
        f[4 * i + 1] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7651)) - 8388736.0f;
        f[4 * i + 2] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7652)) - 8388736.0f;
        f[4 * i + 3] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7653)) - 8388736.0f;
    }
}

constexpr int kAttnQ8Warps = 4;
constexpr int kAttnQ8Unroll = 4;

// One CTA of 4 warps per (kv head, split).  A lane owns 16 consecutive dims of a row (one 16-byte load of K, one of V), so a warp
// reads 32 / (HD / 16) whole rows per instruction and a score costs log2(HD / 16) shuffles per head instead of 5 per position;
// every (warp, row group) is an independent online-softmax stream over every STRIDE-th position of the split (no communication
// inside the loop, kAttnQ8Unroll positions of loads in flight per stream), merged at the end: across the row groups by shuffles,
// across the warps through shared memory.  HG <= 4 query heads per warp: a group of 8 heads is two head groups on two warp pairs
// (both read the same rows; the second read hits L1 / L2).  The row scale is applied once per score (K) and folded into the
// softmax weight (V).  grid (n_splits, n_kv), 128 threads.
template <int HD, int HG>
__global__ void __launch_bounds__(kAttnQ8Warps * 32, 2) attn_q8_split_kernel(const AttnQ8Params p) {
    constexpr int LPR = HD / 16, RPW = 32 / LPR, U = kAttnQ8Unroll;
    __shared__ float sm[kAttnQ8Warps][HG][HD + 2];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, split = blockIdx.x, kh = blockIdx.y, G = p.G;
    const int NHG = (G + HG - 1) / HG, WPG = kAttnQ8Warps / NHG;   // head groups (1 or 2), warps per head group
    const int hgp = warp % NHG, wsub = warp / NHG, sub = lane % LPR, rg = lane / LPR;
    const int kv_len = *p.pos + 1;
    const int chunk = (kv_len + p.n_splits - 1) / p.n_splits;
    const int p_begin = split * chunk, p_end = min(kv_len, p_begin + chunk);
    const int stride = WPG * RPW;
    float q[HG][16], acc[HG][16], m[HG], l[HG];
#pragma unroll
    for (int g = 0; g < HG; g++) {
        const int head = hgp * HG + g;
        m[g] = -INFINITY;
        l[g] = 0.0f;
#pragma unroll
        for (int v = 0; v < 16; v += 4) {
            const float4 t = head < G ? *reinterpret_cast<const float4*>(p.q + (size_t)(kh * G + head) * HD + sub * 16 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            q[g][v] = t.x; q[g][v + 1] = t.y; q[g][v + 2] = t.z; q[g][v + 3] = t.w;
            acc[g][v] = acc[g][v + 1] = acc[g][v + 2] = acc[g][v + 3] = 0.0f;
        }
    }
    const signed char* kb = p.k8 + (size_t)kh * p.max_seq * HD + sub * 16;
    const signed char* vb = p.v8 + (size_t)kh * p.max_seq * HD + sub * 16;
    const float* ksc = p.k_scale + (size_t)kh * p.max_seq;
    const float* vsc = p.v_scale + (size_t)kh * p.max_seq;
    if (p_begin < p_end) {
        for (int base = p_begin + wsub * RPW; base < p_end; base += U * stride) {   // warp-uniform trip count (shuffles inside)
            const int pos0 = base + rg;
            int4 kr[U], vr[U];
            float ks[U], vs[U];
#pragma unroll
            for (int u = 0; u < U; u++) {   // every load of the step first (clamped: stays in range, masked below)
                const int pc = min(pos0 + u * stride, p_end - 1);
                kr[u] = *reinterpret_cast<const int4*>(kb + (size_t)pc * HD);
                vr[u] = *reinterpret_cast<const int4*>(vb + (size_t)pc * HD);
                ks[u] = ksc[pc];
                vs[u] = vsc[pc];
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool valid = pos0 + u * stride < p_end;
                float f[16], s[HG];
                kv_q8_unpack16(kr[u], f);
#pragma unroll
                for (int g = 0; g < HG; g++) {
                    float d = 0.0f;
#pragma unroll
                    for (int v = 0; v < 16; v++) d = fmaf(q[g][v], f[v], d);
                    s[g] = d;
                }
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
                    for (int g = 0; g < HG; g++) s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
                kv_q8_unpack16(vr[u], f);
#pragma unroll
                for (int g = 0; g < HG; g++) {
                    const float sg = valid ? s[g] * ks[u] * p.scale : -INFINITY;
                    const float mn = fmaxf(m[g], sg);
                    const float corr = (m[g] == -INFINITY) ? 0.0f : expf(m[g] - mn);
                    const float w = valid ? expf(sg - mn) : 0.0f;
                    const float wv = w * vs[u];
                    l[g] = l[g] * corr + w;
#pragma unroll
                    for (int v = 0; v < 16; v++) acc[g][v] = fmaf(wv, f[v], acc[g][v] * corr);
                    m[g] = mn;
                }
            }
        }
    }
    // merge the RPW row-group streams of the warp (lanes with the same dims), then the WPG warps of the head group
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
        for (int g = 0; g < HG; g++) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m[g], o), l2 = __shfl_xor_sync(0xffffffffu, l[g], o);
            const float mn = fmaxf(m[g], m2);
            const float ca = (m[g] == -INFINITY) ? 0.0f : expf(m[g] - mn), cb = (m2 == -INFINITY) ? 0.0f : expf(m2 - mn);
            l[g] = l[g] * ca + l2 * cb;
#pragma unroll
            for (int v = 0; v < 16; v++) acc[g][v] = acc[g][v] * ca + __shfl_xor_sync(0xffffffffu, acc[g][v], o) * cb;
            m[g] = mn;
        }
    }
    if (rg == 0) {
#pragma unroll
        for (int g = 0; g < HG; g++) {
#pragma unroll
            for (int v = 0; v < 16; v++) sm[warp][g][sub * 16 + v] = acc[g][v];
            if (sub == 0) { sm[warp][g][HD] = m[g]; sm[warp][g][HD + 1] = l[g]; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G * (HD + 1); i += blockDim.x) {   // element d < HD of head `head`, or (d == HD) its (m, l) pair
        const int head = i / (HD + 1), d = i - head * (HD + 1);
        const int hg2 = head / HG, g = head - hg2 * HG;
        float mm = -INFINITY;
        for (int w = 0; w < WPG; w++) mm = fmaxf(mm, sm[w * NHG + hg2][g][HD]);
        float a = 0.0f, ll = 0.0f;
        for (int w = 0; w < WPG; w++) {
            const float* src = sm[w * NHG + hg2][g];
            if (src[HD] == -INFINITY) continue;
            const float c = expf(src[HD] - mm);
            ll += src[HD + 1] * c;
            a += src[d < HD ? d : 0] * c;
        }
        float* dst = p.part + (((size_t)kh * p.n_splits + split) * G + head) * (HD + 2);
        if (d < HD) dst[d] = a;
        else { dst[HD] = mm; dst[HD + 1] = ll; }
    }
}

// out[head][:] = sum_s acc_s e^(m_s - m) / sum_s l_s e^(m_s - m), splits in order.  grid n_heads, hd threads.  The (m, l) pairs of
// the splits are fetched by one thread each and the coefficients shared, so the loop over splits has no dependent load in it.
constexpr int kAttnQ8MaxSplits = 128;
__global__ void attn_q8_merge_kernel(const AttnQ8Params p, int hd) {
    __shared__ float sc[kAttnQ8MaxSplits], sm_m[kAttnQ8MaxSplits], sm_l[kAttnQ8MaxSplits];
    __shared__ float s_m, s_l;
    pdl_launch_dependents();
    pdl_wait();
    const int head = blockIdx.x, kh = head / p.G, g = head - kh * p.G, d = threadIdx.x;
    const float* base = p.part + (((size_t)kh * p.n_splits) * p.G + g) * (hd + 2);
    const size_t stride = (size_t)p.G * (hd + 2);
    for (int s = d; s < p.n_splits; s += blockDim.x) { sm_m[s] = base[s * stride + hd]; sm_l[s] = base[s * stride + hd + 1]; }
    __syncthreads();
    if (d == 0) {
        float m = -INFINITY;
        for (int s = 0; s < p.n_splits; s++) m = fmaxf(m, sm_m[s]);
        s_m = m;
    }
    __syncthreads();
    for (int s = d; s < p.n_splits; s += blockDim.x) sc[s] = (sm_m[s] == -INFINITY) ? 0.0f : expf(sm_m[s] - s_m);
    __syncthreads();
    if (d == 0) {
        float l = 0.0f;
        for (int s = 0; s < p.n_splits; s++) l += sc[s] * sm_l[s];
        s_l = l;
    }
    float a = 0.0f;
#pragma unroll 8
    for (int s = 0; s < p.n_splits; s++) a = fmaf(base[s * stride + d], sc[s], a);
    __syncthreads();
    p.out[(size_t)head * hd + d] = a / s_l;
}

}  // namespace b200
