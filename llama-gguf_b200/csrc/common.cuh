// common.cuh — shared device helpers for the cuda-b200 backend (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kWarp = 32;

// ggml type ids (reference: src/gguf/constants.rs:56-89)
enum : int { T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q5_0 = 6, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };

__host__ __device__ inline int type_block_elems(int t) {
    switch (t) {
        case T_F32: case T_F16: return 1;
        case T_Q4_0: case T_Q5_0: case T_Q8_0: return 32;
        case T_Q4_K: case T_Q5_K: case T_Q6_K: return 256;
    }
    return 0;
}
__host__ __device__ inline int type_block_bytes(int t) {
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

__device__ __forceinline__ float half_bits_to_float(uint32_t h) {
    return __half2float(__ushort_as_half((unsigned short)(h & 0xffffu)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// f32 -> fp16 operand of the tcgen05 dequant-GEMM path, saturating: a finite activation above 65504 (Qwen2 FFN intermediates can get
// there) must not become inf and then NaN in the residual stream.  NaN stays NaN.
__device__ __forceinline__ __half f2h_sat(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }

// Streaming (read-once) weight loads: non-coherent path, do not allocate in L1.
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// 16-bit pieces of one sector are fetched by neighbouring instructions: keep them in L1.
__device__ __forceinline__ uint32_t ldg_u16(const void* p) {
    unsigned short r;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return (uint32_t)r;
}
__device__ __forceinline__ int ldg_s8(const void* p) { return (int)__ldg((const signed char*)p); }

// Programmatic dependent launch (PDL): every kernel of the decode chain calls
// pdl_launch_dependents() first (the next kernel may start its weight prefetch)
// and pdl_wait() before touching anything a predecessor wrote.  Both are no-ops
// when the kernel was launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 prefetch of a byte range (used before pdl_wait: weights never depend on a predecessor).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// x-vector layout in shared memory: 4 floats of padding after every 64, which makes the
// float4 reads of the Q4_K/Q5_K lane mapping conflict-free.
__host__ __device__ __forceinline__ int xidx(int e) { return e + ((e >> 6) << 2); }
__host__ __device__ __forceinline__ int xpad_floats(int k) { return k + ((k + 63) >> 6) * 4 + 4; }

}  // namespace b200
