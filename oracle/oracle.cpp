// oracle.cpp — CPU restatement of llama-gguf's quantized-forward hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under llama-gguf_b200/ may link, import or
// call this file; it is the *checker* for the CUDA path (tests/, smoke(),
// bench.py's cpu_baseline / --impl reference leg), never the product.
//
// The reference (Lexmata/llama-gguf v0.14.0) is pure Rust and no Rust toolchain
// exists in the build image, so the reference cannot be compiled here
// (DESIGN.md §oracle).  Every function below follows one reference function,
// cited as file:line relative to the reference root, and reproduces its
// operation ORDER (Rust never contracts a*b+c into an FMA: build with
// -ffp-contract=off; FMAs appear only where the reference calls _mm*_fmadd_ps).
//
// Parity pinning: block dequantisation is cross-checked bit-for-bit against
// gguf-py's numpy dequantize (tests/golden/, tests/test_oracle_golden.py) and
// against the reference's own known-answer tests restated in
// tests/test_oracle_kat.py.  The K-quant fused dots and the end-to-end forward
// have NO golden vectors in the reference (SURVEY.md §8c "not pinned"), so for
// those this restatement is the authority: "parity unpinned" beyond dequant.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include <immintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

// ggml type ids carried across every ABI in this repo (gguf/constants.rs:56-89)
enum { T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q5_0 = 6, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };

// ---------------------------------------------------------------------------
// half <-> float, IEEE exact (half 2.7.1: f16::to_f32 / f16::from_f32, RNE)
// ---------------------------------------------------------------------------
static inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline uint16_t f2h(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) {  // inf / nan
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x200u | ((ax >> 13) & 0x3ffu)) : 0));
    }
    if (ax >= 0x477ff000u) {  // rounds to >= 65520 -> inf
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) {  // subnormal half or zero
        if (ax < 0x33000000u) return (uint16_t)sign;  // < 2^-25 -> 0
        int e = (int)(ax >> 23);
        uint32_t man = (ax & 0x7fffffu) | 0x800000u;
        int shift = 126 - e;  // 14..24
        uint32_t hm = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t half_ = 1u << (shift - 1);
        if (rem > half_ || (rem == half_ && (hm & 1))) hm++;
        return (uint16_t)(sign | hm);
    }
    uint32_t e = (ax >> 23) - 112;
    uint32_t man = ax & 0x7fffffu;
    uint32_t hm = man >> 13;
    uint32_t rem = man & 0x1fffu;
    uint16_t h = (uint16_t)((e << 10) | hm);
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) h++;
    return (uint16_t)(sign | h);
}

ORC_API float orc_f16_to_f32(uint16_t h) { return h2f(h); }
ORC_API uint16_t orc_f32_to_f16(float f) { return f2h(f); }

// ---------------------------------------------------------------------------
// Block layouts (tensor/quant/blocks.rs:8-18, 33-44, 60-70, 114-155) as byte
// offsets; #[repr(C)] with 2-byte alignment, so no padding anywhere.
// ---------------------------------------------------------------------------
static inline int block_elems(int t) {
    switch (t) {
        case T_F32: case T_F16: return 1;
        case T_Q4_0: case T_Q5_0: case T_Q8_0: return 32;
        case T_Q4_K: case T_Q5_K: case T_Q6_K: return 256;
    }
    return 0;
}
static inline int block_bytes(int t) {
    switch (t) {
        case T_F32: return 4;
        case T_F16: return 2;
        case T_Q4_0: return 18;
        case T_Q5_0: return 22;
        case T_Q8_0: return 34;
        case T_Q4_K: return 144;
        case T_Q5_K: return 176;
        case T_Q6_K: return 210;
    }
    return 0;
}
ORC_API int orc_block_elems(int t) { return block_elems(t); }
ORC_API int orc_block_bytes(int t) { return block_bytes(t); }

static inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// get_scale_min_k4 unpack shared by Q4_K and Q5_K (dequant.rs:213-225, 270-282)
static inline void unpack_k4(const uint8_t* s, uint8_t* scales, uint8_t* mins) {
    for (int j = 0; j < 4; j++) {
        scales[j] = s[j] & 0x3F;
        mins[j] = s[j + 4] & 0x3F;
    }
    for (int j = 4; j < 8; j++) {
        scales[j] = (uint8_t)((s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4));
        mins[j] = (uint8_t)(((s[j + 4] >> 4) & 0x0F) | ((s[j] >> 6) << 4));
    }
}

// dequant.rs:16-29
static void dequantize_q4_0(const uint8_t* b, float* out) {
    float d = h2f(rd16(b));
    const uint8_t* qs = b + 2;
    for (int i = 0; i < 16; i++) {
        int lo = (int)(qs[i] & 0x0F) - 8;
        int hi = (int)((qs[i] >> 4) & 0x0F) - 8;
        out[i] = (float)lo * d;
        out[i + 16] = (float)hi * d;
    }
}
// dequant.rs:53-74
static void dequantize_q5_0(const uint8_t* b, float* out) {
    float d = h2f(rd16(b));
    uint32_t qh = (uint32_t)b[2] | ((uint32_t)b[3] << 8) | ((uint32_t)b[4] << 16) | ((uint32_t)b[5] << 24);
    const uint8_t* qs = b + 6;
    for (int i = 0; i < 16; i++) {
        int lo4 = qs[i] & 0x0F;
        int hi4 = (qs[i] >> 4) & 0x0F;
        int lo5 = (qh >> i) & 1;
        int hi5 = (qh >> (i + 16)) & 1;
        int lo = (lo4 | (lo5 << 4)) - 16;
        int hi = (hi4 | (hi5 << 4)) - 16;
        out[i] = (float)lo * d;
        out[i + 16] = (float)hi * d;
    }
}
// dequant.rs:103-109
static void dequantize_q8_0(const uint8_t* b, float* out) {
    float d = h2f(rd16(b));
    const int8_t* qs = (const int8_t*)(b + 2);
    for (int i = 0; i < 32; i++) out[i] = (float)qs[i] * d;
}
// dequant.rs:205-256
static void dequantize_q4_k(const uint8_t* b, float* out) {
    float d = h2f(rd16(b));
    float dmin = h2f(rd16(b + 2));
    uint8_t scales[8], mins[8];
    unpack_k4(b + 4, scales, mins);
    const uint8_t* qs = b + 16;
    int o = 0, qp = 0, is = 0;
    for (int g = 0; g < 4; g++) {
        float d1 = d * (float)scales[is];
        float m1 = dmin * (float)mins[is];
        float d2 = d * (float)scales[is + 1];
        float m2 = dmin * (float)mins[is + 1];
        for (int l = 0; l < 32; l++) {
            float q = (float)(qs[qp + l] & 0x0F);
            out[o++] = d1 * q - m1;
        }
        for (int l = 0; l < 32; l++) {
            float q = (float)((qs[qp + l] >> 4) & 0x0F);
            out[o++] = d2 * q - m2;
        }
        qp += 32;
        is += 2;
    }
}
// dequant.rs:262-315
static void dequantize_q5_k(const uint8_t* b, float* out) {
    float d = h2f(rd16(b));
    float dmin = h2f(rd16(b + 2));
    uint8_t scales[8], mins[8];
    unpack_k4(b + 4, scales, mins);
    const uint8_t* qh = b + 16;
    const uint8_t* qs = b + 48;
    int o = 0, qp = 0, is = 0;
    uint8_t u1 = 1, u2 = 2;
    for (int g = 0; g < 4; g++) {
        float d1 = d * (float)scales[is];
        float m1 = dmin * (float)mins[is];
        float d2 = d * (float)scales[is + 1];
        float m2 = dmin * (float)mins[is + 1];
        for (int l = 0; l < 32; l++) {
            float lo4 = (float)(qs[qp + l] & 0x0F);
            float hi5 = (qh[l] & u1) ? 16.0f : 0.0f;
            out[o++] = d1 * (lo4 + hi5) - m1;
        }
        for (int l = 0; l < 32; l++) {
            float hi4 = (float)((qs[qp + l] >> 4) & 0x0F);
            float hi5 = (qh[l] & u2) ? 16.0f : 0.0f;
            out[o++] = d2 * (hi4 + hi5) - m2;
        }
        qp += 32;
        is += 2;
        u1 = (uint8_t)(u1 << 2);
        u2 = (uint8_t)(u2 << 2);
    }
}
// Q6_K element extraction shared by dequant (dequant.rs:321-356) and dot (simd.rs:1098-1146)
static inline void q6k_unpack4(const uint8_t* ql, const uint8_t* qh, int ql_base, int qh_base, int l, int* q) {
    q[0] = (int)((ql[ql_base + l] & 0x0F) | ((qh[qh_base + l] & 0x03) << 4)) - 32;
    q[1] = (int)((ql[ql_base + l + 32] & 0x0F) | (((qh[qh_base + l] >> 2) & 0x03) << 4)) - 32;
    q[2] = (int)((ql[ql_base + l] >> 4) | (((qh[qh_base + l] >> 4) & 0x03) << 4)) - 32;
    q[3] = (int)((ql[ql_base + l + 32] >> 4) | (((qh[qh_base + l] >> 6) & 0x03) << 4)) - 32;
}
static void dequantize_q6_k(const uint8_t* b, float* out) {
    const uint8_t* ql = b;
    const uint8_t* qh = b + 128;
    const int8_t* sc = (const int8_t*)(b + 192);
    float d = h2f(rd16(b + 208));
    for (int n = 0; n < 2; n++) {
        int ql_base = n * 64, qh_base = n * 32, sc_base = n * 8, out_base = n * 128;
        for (int l = 0; l < 32; l++) {
            int is = l / 16;
            int q[4];
            q6k_unpack4(ql, qh, ql_base, qh_base, l, q);
            out[out_base + l] = d * (float)sc[sc_base + is] * (float)q[0];
            out[out_base + l + 32] = d * (float)sc[sc_base + is + 2] * (float)q[1];
            out[out_base + l + 64] = d * (float)sc[sc_base + is + 4] * (float)q[2];
            out[out_base + l + 96] = d * (float)sc[sc_base + is + 6] * (float)q[3];
        }
    }
}

static void dequantize_block(int t, const uint8_t* b, float* out) {
    switch (t) {
        case T_Q4_0: dequantize_q4_0(b, out); break;
        case T_Q5_0: dequantize_q5_0(b, out); break;
        case T_Q8_0: dequantize_q8_0(b, out); break;
        case T_Q4_K: dequantize_q4_k(b, out); break;
        case T_Q5_K: dequantize_q5_k(b, out); break;
        case T_Q6_K: dequantize_q6_k(b, out); break;
        case T_F32: memcpy(out, b, 4); break;
        case T_F16: out[0] = h2f(rd16(b)); break;
    }
}

// Backend::dequantize (cpu/ops.rs:576-916): per block, independent -> parallel
ORC_API int orc_dequantize(int t, const void* src, int64_t n_elems, float* out) {
    int be = block_elems(t), bb = block_bytes(t);
    if (be == 0 || n_elems % be != 0) return -1;
    int64_t nb = n_elems / be;
    const uint8_t* p = (const uint8_t*)src;
#pragma omp parallel for schedule(static) if (nb > 4096)
    for (int64_t i = 0; i < nb; i++) dequantize_block(t, p + i * bb, out + i * be);
    return 0;
}

// ---------------------------------------------------------------------------
// Quantizers (dequant.rs:374-397 Q4_0, 429-449 Q8_0, 455-487 Q5_0, 710-806 Q4_K,
// 813-917 Q5_K, 923-999 Q6_K).  Used only to synthesise random-init weights.
// Rust f32::round = half away from zero = roundf; `as u8/i8` saturate.
// ---------------------------------------------------------------------------
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline void wr16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)(v & 0xff); p[1] = (uint8_t)(v >> 8); }

static void quantize_q4_0(const float* in, uint8_t* b) {
    float amax = 0.0f;
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(in[i]));
    float d = amax / 7.0f;
    float id = d != 0.0f ? 1.0f / d : 0.0f;
    for (int i = 0; i < 16; i++) {
        int lo = clampi((int)roundf(in[i] * id), -8, 7) + 8;
        int hi = clampi((int)roundf(in[i + 16] * id), -8, 7) + 8;
        b[2 + i] = (uint8_t)(lo | (hi << 4));
    }
    wr16(b, f2h(d));
}
static void quantize_q8_0(const float* in, uint8_t* b) {
    float amax = 0.0f;
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(in[i]));
    float d = amax / 127.0f;
    float id = d != 0.0f ? 1.0f / d : 0.0f;
    for (int i = 0; i < 32; i++) b[2 + i] = (uint8_t)(int8_t)clampf(roundf(in[i] * id), -127.0f, 127.0f);
    wr16(b, f2h(d));
}
static void quantize_q5_0(const float* in, uint8_t* b) {
    float amax = 0.0f;
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(in[i]));
    float d = amax / 15.0f;
    float id = d != 0.0f ? 1.0f / d : 0.0f;
    uint32_t qh = 0;
    for (int i = 0; i < 16; i++) {
        int lo = clampi((int)roundf(in[i] * id), -16, 15) + 16;
        int hi = clampi((int)roundf(in[i + 16] * id), -16, 15) + 16;
        b[6 + i] = (uint8_t)((lo & 0x0F) | ((hi & 0x0F) << 4));
        qh |= (uint32_t)((lo >> 4) & 1) << i;
        qh |= (uint32_t)((hi >> 4) & 1) << (i + 16);
    }
    wr16(b, f2h(d));
    b[2] = (uint8_t)qh; b[3] = (uint8_t)(qh >> 8); b[4] = (uint8_t)(qh >> 16); b[5] = (uint8_t)(qh >> 24);
}
// shared front half of quantize_q4_k / quantize_q5_k; qmax = 15 or 31
static void quantize_k45(const float* in, uint8_t* b, int qmax, bool q5) {
    float ranges[8], gmins[8];
    for (int is = 0; is < 8; is++) {
        float gmin = INFINITY, gmax = -INFINITY;
        for (int l = 0; l < 32; l++) {
            gmin = fminf(gmin, in[is * 32 + l]);
            gmax = fmaxf(gmax, in[is * 32 + l]);
        }
        ranges[is] = fmaxf(gmax - gmin, 0.0f);
        gmins[is] = gmin;
    }
    float max_range = 0.0f, max_neg_min = 0.0f;
    for (int is = 0; is < 8; is++) {
        max_range = fmaxf(max_range, ranges[is]);
        max_neg_min = fmaxf(max_neg_min, fmaxf(-gmins[is], 0.0f));
    }
    float d = max_range > 0.0f ? max_range / ((float)qmax * 63.0f) : 1.0f;
    float dmin = max_neg_min > 0.0f ? max_neg_min / 63.0f : 1.0f;
    uint8_t scales[8], mins[8];
    uint8_t* qh = b + 16;
    uint8_t* qs = q5 ? b + 48 : b + 16;
    memset(qs, 0, 128);
    if (q5) memset(qh, 0, 32);
    for (int is = 0; is < 8; is++) {
        uint8_t scale = d > 0.0f ? (uint8_t)clampf(roundf(ranges[is] / ((float)qmax * d)), 0.0f, 63.0f) : 0;
        uint8_t floor1 = ranges[is] > 0.0f ? 1 : 0;
        if (scale < floor1) scale = floor1;
        uint8_t min_val = dmin > 0.0f ? (uint8_t)clampf(roundf(fmaxf(-gmins[is], 0.0f) / dmin), 0.0f, 63.0f) : 0;
        scales[is] = scale;
        mins[is] = min_val;
        float d_scale = d * (float)scale;
        float m = dmin * (float)min_val;
        float id_scale = d_scale > 0.0f ? 1.0f / d_scale : 0.0f;
        int qp = (is / 2) * 32;
        bool high = (is % 2) == 1;
        for (int l = 0; l < 32; l++) {
            uint8_t q = (uint8_t)clampf(roundf((in[is * 32 + l] + m) * id_scale), 0.0f, (float)qmax);
            uint8_t lo4 = q & 0x0F;
            if (high) qs[qp + l] = (uint8_t)((qs[qp + l] & 0x0F) | (lo4 << 4));
            else qs[qp + l] = (uint8_t)((qs[qp + l] & 0xF0) | lo4);
            if (q5 && ((q >> 4) & 1)) qh[l] |= (uint8_t)(1u << is);
        }
    }
    uint8_t* sb = b + 4;
    for (int j = 0; j < 4; j++) {
        sb[j] = (uint8_t)((scales[j] & 0x3F) | ((scales[j + 4] & 0x03) << 6));
        sb[j + 4] = (uint8_t)((mins[j] & 0x3F) | ((mins[j + 4] & 0x03) << 6));
        sb[j + 8] = (uint8_t)(((scales[j + 4] >> 2) & 0x0F) | (((mins[j + 4] >> 2) & 0x0F) << 4));
    }
    wr16(b, f2h(d));
    wr16(b + 2, f2h(dmin));
}
static void quantize_q6_k(const float* in, uint8_t* b) {
    float amax = 0.0f;
    for (int i = 0; i < 256; i++) amax = fmaxf(amax, fabsf(in[i]));
    float d = fmaxf(amax / 16.0f, 1e-10f);
    uint8_t* ql = b;
    uint8_t* qh = b + 128;
    int8_t* scales = (int8_t*)(b + 192);
    for (int n = 0; n < 2; n++) {
        int ql_base = n * 64, qh_base = n * 32, sc_base = n * 8, out_base = n * 128;
        for (int s = 0; s < 8; s++) {
            float gmax = 0.0f;
            for (int i = 0; i < 16; i++) gmax = fmaxf(gmax, fabsf(in[out_base + s * 16 + i]));
            float scale_f = gmax > 1e-10f ? gmax / 31.0f / d : 0.0f;
            int8_t sc = (int8_t)(int)roundf(clampf(scale_f, -128.0f, 127.0f));
            scales[sc_base + s] = (sc == 0 && gmax > 1e-10f) ? 1 : sc;
        }
        for (int l = 0; l < 32; l++) {
            int is = l / 16;
            int q[4];
            for (int j = 0; j < 4; j++) {
                float sc = d * (float)scales[sc_base + is + 2 * j];
                float id = sc != 0.0f ? 1.0f / sc : 0.0f;
                int v = (int)clampf(roundf(in[out_base + l + 32 * j] * id), -32.0f, 31.0f) + 32;
                q[j] = clampi(v, 0, 63);
            }
            ql[ql_base + l] = (uint8_t)((q[0] & 0x0F) | ((q[2] & 0x0F) << 4));
            ql[ql_base + l + 32] = (uint8_t)((q[1] & 0x0F) | ((q[3] & 0x0F) << 4));
            qh[qh_base + l] = (uint8_t)((q[0] >> 4) | ((q[1] >> 4) << 2) | ((q[2] >> 4) << 4) | ((q[3] >> 4) << 6));
        }
    }
    wr16(b + 208, f2h(d));
}

ORC_API int orc_quantize(int t, const float* in, int64_t n_elems, void* out) {
    int be = block_elems(t), bb = block_bytes(t);
    if (be == 0 || n_elems % be != 0) return -1;
    int64_t nb = n_elems / be;
    uint8_t* o = (uint8_t*)out;
    if (t == T_F32) { memcpy(out, in, (size_t)n_elems * 4); return 0; }
    if (t == T_F16) { for (int64_t i = 0; i < n_elems; i++) wr16(o + 2 * i, f2h(in[i])); return 0; }
#pragma omp parallel for schedule(static) if (nb > 1024)
    for (int64_t i = 0; i < nb; i++) {
        const float* x = in + i * be;
        uint8_t* b = o + i * bb;
        switch (t) {
            case T_Q4_0: quantize_q4_0(x, b); break;
            case T_Q5_0: quantize_q5_0(x, b); break;
            case T_Q8_0: quantize_q8_0(x, b); break;
            case T_Q4_K: quantize_k45(x, b, 15, false); break;
            case T_Q5_K: quantize_k45(x, b, 31, true); break;
            case T_Q6_K: quantize_q6_k(x, b); break;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// f32 SIMD helpers with the reference's runtime dispatch (cpu/simd.rs:80-101:
// AVX-512 first, then AVX2, then scalar).
// ---------------------------------------------------------------------------
static int g_has_avx512 = -1, g_has_avx2 = -1;
static inline void detect() {
    if (g_has_avx2 < 0) {
        __builtin_cpu_init();
        g_has_avx2 = (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) ? 1 : 0;
        g_has_avx512 = __builtin_cpu_supports("avx512f") ? 1 : 0;
        const char* e = getenv("ORC_NO_AVX512");
        if (e && e[0] == '1') g_has_avx512 = 0;
    }
}
ORC_API int orc_has_avx512(void) { detect(); return g_has_avx512; }
ORC_API int orc_has_avx2(void) { detect(); return g_has_avx2; }

// simd.rs:199-214
__attribute__((target("avx2"))) static inline float hsum_avx2(__m256 v) {
    __m128 high = _mm256_extractf128_ps(v, 1);
    __m128 low = _mm256_castps256_ps128(v);
    __m128 sum128 = _mm_add_ps(high, low);
    __m128 shuf = _mm_movehdup_ps(sum128);
    __m128 sum64 = _mm_add_ps(sum128, shuf);
    __m128 shuf2 = _mm_movehl_ps(sum64, sum64);
    __m128 sum32 = _mm_add_ss(sum64, shuf2);
    return _mm_cvtss_f32(sum32);
}
// simd.rs:110-136
__attribute__((target("avx2,fma"))) static float dot_f32_avx2(const float* a, const float* b, int64_t n) {
    int64_t chunks = n / 8;
    __m256 sum = _mm256_setzero_ps();
    for (int64_t i = 0; i < chunks; i++)
        sum = _mm256_fmadd_ps(_mm256_loadu_ps(a + i * 8), _mm256_loadu_ps(b + i * 8), sum);
    float r = hsum_avx2(sum);
    for (int64_t i = chunks * 8; i < n; i++) r += a[i] * b[i];
    return r;
}
// simd.rs:140-165.  _mm512_reduce_add_ps order is compiler-defined in both
// toolchains: attention / router dots are tolerance-level, not bit-level.
__attribute__((target("avx512f"))) static float dot_f32_avx512(const float* a, const float* b, int64_t n) {
    int64_t chunks = n / 16;
    __m512 sum = _mm512_setzero_ps();
    for (int64_t i = 0; i < chunks; i++)
        sum = _mm512_fmadd_ps(_mm512_loadu_ps(a + i * 16), _mm512_loadu_ps(b + i * 16), sum);
    float r = _mm512_reduce_add_ps(sum);
    for (int64_t i = chunks * 16; i < n; i++) r += a[i] * b[i];
    return r;
}
static float dot_f32(const float* a, const float* b, int64_t n) {
    detect();
    if (g_has_avx512) return dot_f32_avx512(a, b, n);
    if (g_has_avx2) return dot_f32_avx2(a, b, n);
    float s = 0.0f;  // simd.rs:103-105 (iterator sum: sequential)
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
ORC_API float orc_dot_f32(const float* a, const float* b, int64_t n) { return dot_f32(a, b, n); }

// simd.rs:391-426
__attribute__((target("avx2,fma"))) static void axpy_avx2(float alpha, const float* x, float* y, int64_t n) {
    int64_t chunks = n / 8;
    __m256 va = _mm256_set1_ps(alpha);
    for (int64_t i = 0; i < chunks; i++)
        _mm256_storeu_ps(y + i * 8, _mm256_fmadd_ps(va, _mm256_loadu_ps(x + i * 8), _mm256_loadu_ps(y + i * 8)));
    for (int64_t i = chunks * 8; i < n; i++) y[i] += alpha * x[i];
}
__attribute__((target("avx512f"))) static void axpy_avx512(float alpha, const float* x, float* y, int64_t n) {
    int64_t chunks = n / 16;
    __m512 va = _mm512_set1_ps(alpha);
    for (int64_t i = 0; i < chunks; i++)
        _mm512_storeu_ps(y + i * 16, _mm512_fmadd_ps(va, _mm512_loadu_ps(x + i * 16), _mm512_loadu_ps(y + i * 16)));
    for (int64_t i = chunks * 16; i < n; i++) y[i] += alpha * x[i];
}
static void axpy_f32(float alpha, const float* x, float* y, int64_t n) {
    detect();
    if (g_has_avx512) { axpy_avx512(alpha, x, y, n); return; }
    if (g_has_avx2) { axpy_avx2(alpha, x, y, n); return; }
    for (int64_t i = 0; i < n; i++) y[i] += alpha * x[i];
}
// simd.rs:531-562
__attribute__((target("avx2"))) static float max_f32_avx2(const float* a, int64_t n) {
    if (n == 0) return -INFINITY;
    int64_t chunks = n / 8;
    __m256 vmax = _mm256_set1_ps(-INFINITY);
    for (int64_t i = 0; i < chunks; i++) vmax = _mm256_max_ps(vmax, _mm256_loadu_ps(a + i * 8));
    __m128 high = _mm256_extractf128_ps(vmax, 1);
    __m128 low = _mm256_castps256_ps128(vmax);
    __m128 m128 = _mm_max_ps(high, low);
    __m128 shuf = _mm_movehdup_ps(m128);
    __m128 m64 = _mm_max_ps(m128, shuf);
    __m128 shuf2 = _mm_movehl_ps(m64, m64);
    __m128 m32 = _mm_max_ss(m64, shuf2);
    float r = _mm_cvtss_f32(m32);
    for (int64_t i = chunks * 8; i < n; i++) r = fmaxf(r, a[i]);
    return r;
}
static float max_f32(const float* a, int64_t n) {
    detect();
    if (g_has_avx2) return max_f32_avx2(a, n);
    float r = -INFINITY;
    for (int64_t i = 0; i < n; i++) r = fmaxf(r, a[i]);
    return r;
}
// simd.rs:679-750: max, exp(x-max) with a sequential sum, then *(1/sum)
static void softmax_inplace(float* x, int64_t n) {
    if (n == 0) return;
    float mx = max_f32(x, n);
    float sum = 0.0f;
    for (int64_t i = 0; i < n; i++) {
        x[i] = expf(x[i] - mx);
        sum += x[i];
    }
    float inv = 1.0f / sum;
    for (int64_t i = 0; i < n; i++) x[i] *= inv;
}
ORC_API void orc_softmax(const float* x, float* out, int64_t n) {
    if (out != x) memcpy(out, x, (size_t)n * 4);
    softmax_inplace(out, n);
}
// simd.rs:785-823
__attribute__((target("avx2,fma"))) static float sumsq_avx2(const float* x, int64_t n) {
    int64_t chunks = n / 8;
    __m256 sum = _mm256_setzero_ps();
    for (int64_t i = 0; i < chunks; i++) {
        __m256 v = _mm256_loadu_ps(x + i * 8);
        sum = _mm256_fmadd_ps(v, v, sum);
    }
    float r = hsum_avx2(sum);
    for (int64_t i = chunks * 8; i < n; i++) r += x[i] * x[i];
    return r;
}
static float sum_of_squares(const float* x, int64_t n) {
    detect();
    if (g_has_avx2) return sumsq_avx2(x, n);
    float s = 0.0f;
    for (int64_t i = 0; i < n; i++) s += x[i] * x[i];
    return s;
}
// simd.rs:847-899: inv = 1/sqrt(ss/n + eps); out = (x*inv)*w (two multiplies)
static void rms_norm(const float* x, const float* w, float eps, float* out, int64_t n) {
    float ss = sum_of_squares(x, n);
    float rms = sqrtf(ss / (float)n + eps);
    float inv = 1.0f / rms;
    for (int64_t i = 0; i < n; i++) {
        float scaled = x[i] * inv;
        out[i] = scaled * w[i];
    }
}
// Backend::rms_norm (cpu/ops.rs:392-422): row-wise over the last dimension
ORC_API void orc_rms_norm(const float* x, const float* w, float eps, float* out, int64_t n_rows, int64_t hidden) {
    for (int64_t r = 0; r < n_rows; r++) rms_norm(x + r * hidden, w, eps, out + r * hidden, hidden);
}
// cpu/ops.rs:303-325 and simd.rs:598-649
static inline float silu1(float x) { return x / (1.0f + expf(-x)); }
ORC_API void orc_silu(const float* x, float* out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = silu1(x[i]); }
ORC_API void orc_silu_mul_inplace(float* gate, const float* up, int64_t n) {
    for (int64_t i = 0; i < n; i++) gate[i] = silu1(gate[i]) * up[i];
}
ORC_API void orc_add(const float* a, const float* b, float* o, int64_t n) { for (int64_t i = 0; i < n; i++) o[i] = a[i] + b[i]; }
ORC_API void orc_mul(const float* a, const float* b, float* o, int64_t n) { for (int64_t i = 0; i < n; i++) o[i] = a[i] * b[i]; }
ORC_API void orc_scale(const float* a, float s, float* o, int64_t n) { for (int64_t i = 0; i < n; i++) o[i] = a[i] * s; }

// ---------------------------------------------------------------------------
// Fused quantised dot products (cpu/simd.rs:931-1146): scalar, strictly
// sequential f32 — the reference's CPU inner loop.
// ---------------------------------------------------------------------------
static float dot_q4_0(const uint8_t* w, int64_t nb, const float* x) {
    float sum = 0.0f;
    int64_t off = 0;
    for (int64_t b = 0; b < nb; b++, w += 18) {
        float d = h2f(rd16(w));
        float acc_lo = 0.0f, acc_hi = 0.0f;
        for (int i = 0; i < 16; i++) {
            uint8_t byte = w[2 + i];
            acc_lo += (float)((int)(byte & 0x0F) - 8) * x[off + i];
            acc_hi += (float)((int)((byte >> 4) & 0x0F) - 8) * x[off + i + 16];
        }
        sum += d * (acc_lo + acc_hi);
        off += 32;
    }
    return sum;
}
static float dot_q8_0(const uint8_t* w, int64_t nb, const float* x) {
    float sum = 0.0f;
    int64_t off = 0;
    for (int64_t b = 0; b < nb; b++, w += 34) {
        float d = h2f(rd16(w));
        const int8_t* qs = (const int8_t*)(w + 2);
        float acc = 0.0f;
        for (int i = 0; i < 32; i++) acc += (float)qs[i] * x[off + i];
        sum += d * acc;
        off += 32;
    }
    return sum;
}
static float dot_q4_k(const uint8_t* w, int64_t nb, const float* x) {
    float sum = 0.0f;
    int64_t xo = 0;
    for (int64_t b = 0; b < nb; b++, w += 144) {
        float d = h2f(rd16(w));
        float dmin = h2f(rd16(w + 2));
        uint8_t scales[8], mins[8];
        unpack_k4(w + 4, scales, mins);
        const uint8_t* qs = w + 16;
        int qp = 0, is = 0;
        for (int g = 0; g < 4; g++) {
            float d1 = d * (float)scales[is], m1 = dmin * (float)mins[is];
            float d2 = d * (float)scales[is + 1], m2 = dmin * (float)mins[is + 1];
            float qa = 0.0f, xa = 0.0f;
            for (int l = 0; l < 32; l++) {
                float q = (float)(qs[qp + l] & 0x0F);
                qa += q * x[xo + l];
                xa += x[xo + l];
            }
            sum += d1 * qa - m1 * xa;
            xo += 32;
            qa = 0.0f; xa = 0.0f;
            for (int l = 0; l < 32; l++) {
                float q = (float)((qs[qp + l] >> 4) & 0x0F);
                qa += q * x[xo + l];
                xa += x[xo + l];
            }
            sum += d2 * qa - m2 * xa;
            xo += 32;
            qp += 32;
            is += 2;
        }
    }
    return sum;
}
static float dot_q5_k(const uint8_t* w, int64_t nb, const float* x) {
    float sum = 0.0f;
    int64_t xo = 0;
    for (int64_t b = 0; b < nb; b++, w += 176) {
        float d = h2f(rd16(w));
        float dmin = h2f(rd16(w + 2));
        uint8_t scales[8], mins[8];
        unpack_k4(w + 4, scales, mins);
        const uint8_t* qh = w + 16;
        const uint8_t* qs = w + 48;
        int qp = 0, is = 0;
        uint8_t u1 = 1, u2 = 2;
        for (int g = 0; g < 4; g++) {
            float d1 = d * (float)scales[is], m1 = dmin * (float)mins[is];
            float d2 = d * (float)scales[is + 1], m2 = dmin * (float)mins[is + 1];
            float qa = 0.0f, xa = 0.0f;
            for (int l = 0; l < 32; l++) {
                float lo4 = (float)(qs[qp + l] & 0x0F);
                float hi5 = (qh[l] & u1) ? 16.0f : 0.0f;
                qa += (lo4 + hi5) * x[xo + l];
                xa += x[xo + l];
            }
            sum += d1 * qa - m1 * xa;
            xo += 32;
            qa = 0.0f; xa = 0.0f;
            for (int l = 0; l < 32; l++) {
                float hi4 = (float)((qs[qp + l] >> 4) & 0x0F);
                float hi5 = (qh[l] & u2) ? 16.0f : 0.0f;
                qa += (hi4 + hi5) * x[xo + l];
                xa += x[xo + l];
            }
            sum += d2 * qa - m2 * xa;
            xo += 32;
            qp += 32;
            is += 2;
            u1 = (uint8_t)(u1 << 2);
            u2 = (uint8_t)(u2 << 2);
        }
    }
    return sum;
}
static float dot_q6_k(const uint8_t* w, int64_t nb, const float* x) {
    float sum = 0.0f;
    int64_t xo = 0;
    for (int64_t b = 0; b < nb; b++, w += 210) {
        const uint8_t* ql = w;
        const uint8_t* qh = w + 128;
        const int8_t* sc = (const int8_t*)(w + 192);
        float d = h2f(rd16(w + 208));
        for (int n = 0; n < 2; n++) {
            int ql_base = n * 64, qh_base = n * 32, sc_base = n * 8, ob = n * 128;
            for (int l = 0; l < 32; l++) {
                int is = l / 16;
                int q[4];
                q6k_unpack4(ql, qh, ql_base, qh_base, l, q);
                float s1 = (float)sc[sc_base + is], s2 = (float)sc[sc_base + is + 2];
                float s3 = (float)sc[sc_base + is + 4], s4 = (float)sc[sc_base + is + 6];
                sum += d * s1 * (float)q[0] * x[xo + ob + l];
                sum += d * s2 * (float)q[1] * x[xo + ob + l + 32];
                sum += d * s3 * (float)q[2] * x[xo + ob + l + 64];
                sum += d * s4 * (float)q[3] * x[xo + ob + l + 96];
            }
        }
        xo += 256;
    }
    return sum;
}

// one output row of vec_mat_q.  Types without a fused dot (Q5_0, F16...) take
// fused_vecmat_dispatch's `_ =>` arm (cpu/ops.rs:1182-1189): dequantise, then
// simd::dot_f32(x, row).
static float row_dot(int t, const uint8_t* row, const float* x, int64_t k, float* scratch) {
    switch (t) {
        case T_Q4_0: return dot_q4_0(row, k / 32, x);
        case T_Q8_0: return dot_q8_0(row, k / 32, x);
        case T_Q4_K: return dot_q4_k(row, k / 256, x);
        case T_Q5_K: return dot_q5_k(row, k / 256, x);
        case T_Q6_K: return dot_q6_k(row, k / 256, x);
        default: {
            int be = block_elems(t), bb = block_bytes(t);
            for (int64_t b = 0; b < k / be; b++) dequantize_block(t, row + b * bb, scratch + b * be);
            return dot_f32(x, scratch, k);
        }
    }
}
ORC_API float orc_dot_q(int t, const void* row, const float* x, int64_t k) {
    std::vector<float> scratch((size_t)k);
    return row_dot(t, (const uint8_t*)row, x, k, scratch.data());
}

// Backend::vec_mat_q (cpu/ops.rs:1008-1039 -> 1123-1191): rayon over the n
// output rows; each row is k/bs contiguous blocks.  F32 weights take
// Backend::vec_mat (cpu/ops.rs:959-1005): plain sequential scalar sum.
static void vec_mat_any(int t, const void* w, const float* x, float* out, int64_t k, int64_t n) {
    const uint8_t* base = (const uint8_t*)w;
    if (t == T_F32) {
        const float* wf = (const float*)w;
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < n; j++) {
            float sum = 0.0f;
            for (int64_t i = 0; i < k; i++) sum += x[i] * wf[i + j * k];
            out[j] = sum;
        }
        return;
    }
    int64_t row_bytes = k / block_elems(t) * block_bytes(t);
    bool need_scratch = !(t == T_Q4_0 || t == T_Q8_0 || t == T_Q4_K || t == T_Q5_K || t == T_Q6_K);
#pragma omp parallel
    {
        std::vector<float> scratch(need_scratch ? (size_t)k : 0);
#pragma omp for schedule(static)
        for (int64_t j = 0; j < n; j++) out[j] = row_dot(t, base + j * row_bytes, x, k, scratch.data());
    }
}
ORC_API int orc_vec_mat_q(int t, const void* w, const float* x, float* out, int64_t k, int64_t n) {
    if (block_elems(t) == 0 || k % block_elems(t) != 0) return -1;
    vec_mat_any(t, w, x, out, k, n);
    return 0;
}

// Backend::rope (cpu/ops.rs:1216-1337); seq_len = 1 on the decode path
static void rope_tensor(float* data, int n_heads, int head_dim, int pos, float base, float scale, int neox) {
    int half = head_dim / 2;
    for (int h = 0; h < n_heads; h++) {
        float position = (float)pos / scale;
        float* p = data + (int64_t)h * head_dim;
        for (int i = 0; i < half; i++) {
            float freq = 1.0f / powf(base, (float)(2 * i) / (float)head_dim);
            float theta = position * freq;
            float c = cosf(theta), s = sinf(theta);
            int i0 = neox ? i : 2 * i;
            int i1 = neox ? i + half : 2 * i + 1;
            float x0 = p[i0], x1 = p[i1];
            p[i0] = x0 * c - x1 * s;
            p[i1] = x0 * s + x1 * c;
        }
    }
}
ORC_API void orc_rope(float* q, float* k, int n_heads, int n_kv_heads, int head_dim, int pos, float base, float scale,
                      int neox) {
    rope_tensor(q, n_heads, head_dim, pos, base, scale, neox);
    rope_tensor(k, n_kv_heads, head_dim, pos, base, scale, neox);
}

// Backend::attention_cached (cpu/ops.rs:1479-1537): rayon over heads
ORC_API void orc_attention_cached(const float* q, const float* kc, const float* vc, float* out, int n_heads,
                                  int n_kv_heads, int head_dim, int max_seq, float scale, int kv_len) {
    int qpk = n_heads / n_kv_heads;
    int64_t head_stride = (int64_t)max_seq * head_dim;
#pragma omp parallel
    {
        std::vector<float> scores((size_t)kv_len);
#pragma omp for schedule(static)
        for (int h = 0; h < n_heads; h++) {
            int kvh = h / qpk;
            const float* qv = q + (int64_t)h * head_dim;
            const float* kb = kc + kvh * head_stride;
            const float* vb = vc + kvh * head_stride;
            for (int p = 0; p < kv_len; p++) scores[p] = dot_f32(qv, kb + (int64_t)p * head_dim, head_dim) * scale;
            softmax_inplace(scores.data(), kv_len);
            float* o = out + (int64_t)h * head_dim;
            for (int i = 0; i < head_dim; i++) o[i] = 0.0f;
            for (int p = 0; p < kv_len; p++)
                if (scores[p] > 1e-8f) axpy_f32(scores[p], vb + (int64_t)p * head_dim, o, head_dim);
        }
    }
}

// MoeRouter::route (model/moe.rs:128-198), normalize=false (loader.rs:1155-1159)
ORC_API void orc_moe_route(const float* h, const float* w_router, int hidden, int n_experts, int top_k, int* idx_out,
                           float* w_out) {
    std::vector<std::pair<int, float>> il((size_t)n_experts);
    for (int e = 0; e < n_experts; e++) il[e] = {e, dot_f32(h, w_router + (int64_t)e * hidden, hidden)};
    std::stable_sort(il.begin(), il.end(),
                     [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
    float mx = -INFINITY;
    for (int i = 0; i < top_k; i++) mx = fmaxf(mx, il[i].second);
    float exp_sum = 0.0f;
    for (int i = 0; i < top_k; i++) exp_sum += expf(il[i].second - mx);
    for (int i = 0; i < top_k; i++) {
        idx_out[i] = il[i].first;
        w_out[i] = expf(il[i].second - mx) / exp_sum;
    }
}

// bench argmax (main.rs:1816-1821): max_by -> LAST maximal element wins
ORC_API int orc_argmax_last(const float* v, int64_t n) {
    int64_t best = 0;
    for (int64_t i = 1; i < n; i++)
        if (v[i] >= v[best]) best = i;
    return (int)best;
}

// ---------------------------------------------------------------------------
// Whole-model forward: LlamaModel::forward (model/llama.rs:275-362),
// TransformerLayer::forward serial-residual arm (model/layers.rs:1187-1244),
// Attention::forward (layers.rs:409-704), FeedForward::forward (:908-929),
// MoeLayer::forward (model/moe.rs:321-413).  KV cache [nkv][max_seq][hd] f32
// per layer (model/mod.rs:83-108).
// ---------------------------------------------------------------------------
struct OrcTensor {
    int type = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    int n_dims = 0;
    std::vector<uint8_t> data;
};
struct OrcDesc {  // mirrors include/llama_b200.h:b200_model_desc field for field
    int32_t hidden, n_layers, n_heads, n_kv_heads, head_dim, ffn, vocab, max_seq_len;
    float norm_eps, rope_base, rope_scale;
    int32_t rope_neox, n_experts, n_experts_used, expert_ffn, tied_output, max_batch;
};
struct OrcModel {
    OrcDesc d;
    std::map<std::string, OrcTensor> t;
    std::vector<std::vector<float>> kc, vc;  // per layer
    int position = 0;
    std::vector<float> hiddens;  // [(n_layers+1) * hidden] of the last processed token
    int faithful_embedding = 0;  // 1: dequantise the whole table per call (llama.rs:288)
    int kv_format = 0;           // 1: rows pass through QuantizedKVCache's Int8 format on their way into the cache (kv_quantized.rs)
    const OrcTensor* get(const std::string& n) const {
        auto it = t.find(n);
        return it == t.end() ? nullptr : &it->second;
    }
};

ORC_API OrcModel* orc_model_create(const OrcDesc* d) {
    OrcModel* m = new OrcModel();
    m->d = *d;
    if (m->d.head_dim <= 0) m->d.head_dim = m->d.hidden / m->d.n_heads;
    if (m->d.rope_scale == 0.0f) m->d.rope_scale = 1.0f;
    m->kc.resize(d->n_layers);
    m->vc.resize(d->n_layers);
    size_t kv = (size_t)m->d.n_kv_heads * m->d.max_seq_len * m->d.head_dim;
    for (int l = 0; l < d->n_layers; l++) {
        m->kc[l].assign(kv, 0.0f);
        m->vc[l].assign(kv, 0.0f);
    }
    m->hiddens.assign((size_t)(d->n_layers + 1) * d->hidden, 0.0f);
    return m;
}
ORC_API void orc_model_destroy(OrcModel* m) { delete m; }
ORC_API int orc_model_set_tensor(OrcModel* m, const char* name, int type, const int64_t* ne, int n_dims,
                                 const void* data, int64_t nbytes) {
    OrcTensor& t = m->t[name];
    t.type = type;
    t.n_dims = n_dims;
    for (int i = 0; i < 4; i++) t.ne[i] = i < n_dims ? ne[i] : 1;
    t.data.assign((const uint8_t*)data, (const uint8_t*)data + nbytes);
    return 0;
}
ORC_API void orc_model_reset(OrcModel* m) {
    m->position = 0;
    for (auto& v : m->kc) std::fill(v.begin(), v.end(), 0.0f);
    for (auto& v : m->vc) std::fill(v.begin(), v.end(), 0.0f);
}
ORC_API int orc_model_position(const OrcModel* m) { return m->position; }
ORC_API void orc_model_set_faithful_embedding(OrcModel* m, int on) { m->faithful_embedding = on; }

// quantize_int8 (src/model/kv_quantized.rs:366-387): scale = max|x| / 127 (1.0 for an all-zero row), q = clamp(round(x / scale));
// f32::round is half away from zero = roundf
ORC_API float orc_kv_quantize_int8(const float* x, int n, int8_t* q) {
    float max_abs = 0.0f;
    for (int i = 0; i < n; i++) max_abs = std::max(max_abs, std::fabs(x[i]));
    const float scale = max_abs > 1e-10f ? max_abs / 127.0f : 1.0f;
    for (int i = 0; i < n; i++) {
        float v = roundf(x[i] / scale);
        v = std::min(std::max(v, -128.0f), 127.0f);
        q[i] = (int8_t)v;
    }
    return scale;
}
// dequantize_int8 (kv_quantized.rs:390-392)
ORC_API void orc_kv_dequantize_int8(const int8_t* q, int n, float scale, float* out) {
    for (int i = 0; i < n; i++) out[i] = (float)q[i] * scale;
}
// 0 = F32 KVCache (model/mod.rs:83-108); 1 = QuantizedKVCache Int8: write_kv stores quantize_int8(row) per (head, position)
// (kv_quantized.rs:143-216) and read_k_range / read_v_range return dequantize_int8 of it (:230-310) -- the attention then sees
// q * scale, which is what the f32 cache of this model holds in that mode
ORC_API void orc_model_set_kv_format(OrcModel* m, int format) { m->kv_format = format; }
ORC_API void orc_model_get_hidden(const OrcModel* m, int layer, float* out) {
    memcpy(out, m->hiddens.data() + (size_t)layer * m->d.hidden, (size_t)m->d.hidden * 4);
}
ORC_API void orc_model_get_kv(const OrcModel* m, int layer, int which, float* out) {
    const std::vector<float>& v = which ? m->vc[layer] : m->kc[layer];
    memcpy(out, v.data(), v.size() * 4);
}

// Linear::forward (layers.rs:56-77): vec_mat_q / vec_mat, then bias
static int linear(const OrcModel* m, const std::string& wname, const std::string& bname, const float* x, float* out) {
    const OrcTensor* w = m->get(wname);
    if (!w) { fprintf(stderr, "oracle: missing tensor %s\n", wname.c_str()); return -1; }
    vec_mat_any(w->type, w->data.data(), x, out, w->ne[0], w->ne[1]);
    if (!bname.empty()) {
        const OrcTensor* b = m->get(bname);
        if (b) {
            const float* bd = (const float*)b->data.data();
            for (int64_t i = 0; i < w->ne[1]; i++) out[i] += bd[i];
        }
    }
    return 0;
}

static int layer_forward(OrcModel* m, int l, float* hidden, int pos) {
    const OrcDesc& d = m->d;
    int H = d.hidden, nh = d.n_heads, nkv = d.n_kv_heads, hd = d.head_dim;
    std::string p = "blk." + std::to_string(l) + ".";
    std::vector<float> norm(H), q((size_t)nh * hd), k((size_t)nkv * hd), v((size_t)nkv * hd), attn((size_t)nh * hd),
        ao(H), h(H), fn(H), fo(H);
    const OrcTensor* an = m->get(p + "attn_norm.weight");
    if (!an) return -1;
    rms_norm(hidden, (const float*)an->data.data(), d.norm_eps, norm.data(), H);
    if (linear(m, p + "attn_q.weight", p + "attn_q.bias", norm.data(), q.data())) return -1;
    if (linear(m, p + "attn_k.weight", p + "attn_k.bias", norm.data(), k.data())) return -1;
    if (linear(m, p + "attn_v.weight", p + "attn_v.bias", norm.data(), v.data())) return -1;
    orc_rope(q.data(), k.data(), nh, nkv, hd, pos, d.rope_base, d.rope_scale, d.rope_neox);
    float* kc = m->kc[l].data();
    float* vc = m->vc[l].data();
    for (int hh = 0; hh < nkv; hh++) {  // layers.rs:580-600
        float* kdst = kc + ((int64_t)hh * d.max_seq_len + pos) * hd;
        float* vdst = vc + ((int64_t)hh * d.max_seq_len + pos) * hd;
        if (m->kv_format == 1) {   // QuantizedKVCache::write_kv then read_*_range (kv_quantized.rs:143-216, 230-310)
            std::vector<int8_t> q8((size_t)hd);
            float s = orc_kv_quantize_int8(k.data() + (int64_t)hh * hd, hd, q8.data());
            orc_kv_dequantize_int8(q8.data(), hd, s, kdst);
            s = orc_kv_quantize_int8(v.data() + (int64_t)hh * hd, hd, q8.data());
            orc_kv_dequantize_int8(q8.data(), hd, s, vdst);
            continue;
        }
        memcpy(kdst, k.data() + (int64_t)hh * hd, (size_t)hd * 4);
        memcpy(vdst, v.data() + (int64_t)hh * hd, (size_t)hd * 4);
    }
    float scale = 1.0f / sqrtf((float)hd);  // layers.rs:374
    orc_attention_cached(q.data(), kc, vc, attn.data(), nh, nkv, hd, d.max_seq_len, scale, pos + 1);
    if (linear(m, p + "attn_output.weight", "", attn.data(), ao.data())) return -1;
    for (int i = 0; i < H; i++) h[i] = ao[i] + hidden[i];  // layers.rs:1202-1208 (h = attn_out; h += x)
    const OrcTensor* fnw = m->get(p + "ffn_norm.weight");
    if (!fnw) return -1;
    rms_norm(h.data(), (const float*)fnw->data.data(), d.norm_eps, fn.data(), H);
    if (d.n_experts > 0) {
        const OrcTensor* router = m->get(p + "ffn_gate_inp.weight");
        const OrcTensor* ge = m->get(p + "ffn_gate_exps.weight");
        const OrcTensor* ue = m->get(p + "ffn_up_exps.weight");
        const OrcTensor* de = m->get(p + "ffn_down_exps.weight");
        if (!router || !ge || !ue || !de) return -1;
        int E = d.n_experts, K = d.n_experts_used;
        int I = (int)ge->ne[1];
        std::vector<int> idx(K);
        std::vector<float> wts(K);
        orc_moe_route(fn.data(), (const float*)router->data.data(), H, E, K, idx.data(), wts.data());
        std::fill(fo.begin(), fo.end(), 0.0f);
        std::vector<float> g(I), gs(I), u(I), inter(I), eo(H);
        int64_t gu_bytes = (int64_t)ge->data.size() / E, dn_bytes = (int64_t)de->data.size() / E;
        for (int s = 0; s < K; s++) {  // MoeExpert::forward (moe.rs:227-268)
            int e = idx[s];
            vec_mat_any(ge->type, ge->data.data() + e * gu_bytes, fn.data(), g.data(), H, I);
            orc_silu(g.data(), gs.data(), I);
            vec_mat_any(ue->type, ue->data.data() + e * gu_bytes, fn.data(), u.data(), H, I);
            orc_mul(gs.data(), u.data(), inter.data(), I);
            vec_mat_any(de->type, de->data.data() + e * dn_bytes, inter.data(), eo.data(), I, H);
            for (int i = 0; i < H; i++) fo[i] += wts[s] * eo[i];  // moe.rs:363-368
        }
    } else {
        int I = d.ffn;
        std::vector<float> g(I), u(I);
        if (linear(m, p + "ffn_gate.weight", "", fn.data(), g.data())) return -1;
        if (linear(m, p + "ffn_up.weight", "", fn.data(), u.data())) return -1;
        orc_silu_mul_inplace(g.data(), u.data(), I);
        if (linear(m, p + "ffn_down.weight", "", g.data(), fo.data())) return -1;
    }
    for (int i = 0; i < H; i++) hidden[i] = fo[i] + h[i];  // layers.rs:1235-1241 (ffn_out += h)
    return 0;
}

// embedding row of token t = elements [t*H, (t+1)*H) of dequantised token_embd
static int embed(OrcModel* m, uint32_t tok, float* out) {
    const OrcTensor* e = m->get("token_embd.weight");
    if (!e || (int)tok >= m->d.vocab) return -1;
    int H = m->d.hidden;
    if (m->faithful_embedding && e->type != T_F32) {
        std::vector<float> all((size_t)H * m->d.vocab);
        orc_dequantize(e->type, e->data.data(), (int64_t)H * m->d.vocab, all.data());
        memcpy(out, all.data() + (size_t)tok * H, (size_t)H * 4);
        return 0;
    }
    int64_t row_bytes = (int64_t)H / block_elems(e->type) * block_bytes(e->type);
    return orc_dequantize(e->type, e->data.data() + tok * row_bytes, H, out);
}

// LlamaModel::forward: layer-major, token-minor; logits of the LAST token.
// logits may be NULL (prefill without logits = GpuInference::prefill_token).
ORC_API int orc_model_forward(OrcModel* m, const uint32_t* tokens, int n_tokens, float* logits) {
    const OrcDesc& d = m->d;
    int H = d.hidden;
    if (n_tokens <= 0) return -1;
    if (m->position + n_tokens > d.max_seq_len) return -2;  // ContextLengthExceeded
    std::vector<float> hs((size_t)n_tokens * H);
    for (int t = 0; t < n_tokens; t++)
        if (embed(m, tokens[t], hs.data() + (size_t)t * H)) return -3;
    memcpy(m->hiddens.data(), hs.data() + (size_t)(n_tokens - 1) * H, (size_t)H * 4);
    for (int l = 0; l < d.n_layers; l++) {
        for (int t = 0; t < n_tokens; t++)
            if (layer_forward(m, l, hs.data() + (size_t)t * H, m->position + t)) return -4;
        memcpy(m->hiddens.data() + (size_t)(l + 1) * H, hs.data() + (size_t)(n_tokens - 1) * H, (size_t)H * 4);
    }
    m->position += n_tokens;
    if (!logits) return 0;
    // compute_logits (llama.rs:247-266); tied output if output.weight absent (loader.rs:349-355)
    const OrcTensor* on = m->get("output_norm.weight");
    if (!on) return -5;
    std::vector<float> normed(H);
    rms_norm(hs.data() + (size_t)(n_tokens - 1) * H, (const float*)on->data.data(), d.norm_eps, normed.data(), H);
    const OrcTensor* ow = m->get("output.weight");
    if (!ow) ow = m->get("token_embd.weight");
    if (!ow) return -5;
    vec_mat_any(ow->type, ow->data.data(), normed.data(), logits, ow->ne[0], ow->ne[1]);
    return 0;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
