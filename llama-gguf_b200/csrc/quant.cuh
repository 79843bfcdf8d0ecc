// quant.cuh — GGUF block formats, bit-exact element dequantisation.
//
// Layouts are the reference's #[repr(C)] structs (src/tensor/quant/blocks.rs:8-18,
// 33-44, 60-70, 114-155), kept byte-for-byte as they sit in the GGUF file.  The
// arithmetic follows src/tensor/quant/dequant.rs and is written with __fmul_rn /
// __fsub_rn so nvcc cannot contract `d1*q - m1` into an FMA: outputs are
// bit-identical to the reference (tests/test_gpu_dequant.py).
#pragma once
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t rd_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// get_scale_min_k4 (dequant.rs:213-225): 6-bit scale and min of sub-block j
__device__ __forceinline__ void scale_min_k4(const uint8_t* s, int j, int& sc, int& mn) {
    if (j < 4) {
        sc = s[j] & 0x3F;
        mn = s[j + 4] & 0x3F;
    } else {
        sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4);
        mn = ((s[j + 4] >> 4) & 0x0F) | ((s[j] >> 6) << 4);
    }
}

// One element `e` (0..block_elems) of the block at `b`.
__device__ __forceinline__ float dequant_elem(int type, const uint8_t* b, int e) {
    switch (type) {
        case T_F32: {
            uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
            return __uint_as_float(v);
        }
        case T_F16:
            return half_bits_to_float(rd_u16(b));
        case T_Q4_0: {  // dequant.rs:16-29
            float d = half_bits_to_float(rd_u16(b));
            int i = e & 15;
            int q = (e < 16) ? (b[2 + i] & 0x0F) : ((b[2 + i] >> 4) & 0x0F);
            return __fmul_rn((float)(q - 8), d);
        }
        case T_Q5_0: {  // dequant.rs:53-74
            float d = half_bits_to_float(rd_u16(b));
            uint32_t qh = (uint32_t)b[2] | ((uint32_t)b[3] << 8) | ((uint32_t)b[4] << 16) | ((uint32_t)b[5] << 24);
            int i = e & 15;
            int q4 = (e < 16) ? (b[6 + i] & 0x0F) : ((b[6 + i] >> 4) & 0x0F);
            int q5 = (qh >> e) & 1;  // bit i for the low half, bit i+16 for the high half
            return __fmul_rn((float)((q4 | (q5 << 4)) - 16), d);
        }
        case T_Q8_0: {  // dequant.rs:103-109
            float d = half_bits_to_float(rd_u16(b));
            return __fmul_rn((float)(int)(signed char)b[2 + e], d);
        }
        case T_Q4_K: {  // dequant.rs:205-256
            float d = half_bits_to_float(rd_u16(b));
            float dmin = half_bits_to_float(rd_u16(b + 2));
            int j = e >> 5, l = e & 31, sc, mn;
            scale_min_k4(b + 4, j, sc, mn);
            uint8_t byte = b[16 + (j >> 1) * 32 + l];
            int q = (j & 1) ? (byte >> 4) : (byte & 0x0F);
            float d1 = __fmul_rn(d, (float)sc);
            float m1 = __fmul_rn(dmin, (float)mn);
            return __fsub_rn(__fmul_rn(d1, (float)q), m1);
        }
        case T_Q5_K: {  // dequant.rs:262-315
            float d = half_bits_to_float(rd_u16(b));
            float dmin = half_bits_to_float(rd_u16(b + 2));
            int j = e >> 5, l = e & 31, sc, mn;
            scale_min_k4(b + 4, j, sc, mn);
            uint8_t byte = b[48 + (j >> 1) * 32 + l];
            int q = (j & 1) ? (byte >> 4) : (byte & 0x0F);
            float hi = ((b[16 + l] >> j) & 1) ? 16.0f : 0.0f;
            float d1 = __fmul_rn(d, (float)sc);
            float m1 = __fmul_rn(dmin, (float)mn);
            return __fsub_rn(__fmul_rn(d1, __fadd_rn((float)q, hi)), m1);
        }
        case T_Q6_K: {  // dequant.rs:321-356
            const uint8_t* ql = b;
            const uint8_t* qh = b + 128;
            const signed char* sc = (const signed char*)(b + 192);
            float d = half_bits_to_float(rd_u16(b + 208));
            int n = e >> 7, r = e & 127, c = r >> 5, l = r & 31;
            int is = l >> 4;
            uint8_t lb = ql[n * 64 + l + ((c & 1) ? 32 : 0)];
            int nib = (c & 2) ? (lb >> 4) : (lb & 0x0F);
            int hb = (qh[n * 32 + l] >> (2 * c)) & 3;
            int q = (nib | (hb << 4)) - 32;
            float s = (float)(int)sc[n * 8 + is + 2 * c];
            return __fmul_rn(__fmul_rn(d, s), (float)q);
        }
    }
    return 0.0f;
}

// Backend::dequantize (src/backend/mod.rs; cpu/ops.rs:576-916).  One thread per element.
__global__ void dequantize_kernel(int type, const uint8_t* __restrict__ src, long long n_elems, float* __restrict__ out) {
    const int be = type_block_elems(type), bb = type_block_bytes(type);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (long long)gridDim.x * blockDim.x) {
        long long blk = i / be;
        int e = (int)(i - blk * be);
        out[i] = dequant_elem(type, src + blk * bb, e);
    }
}

}  // namespace b200
