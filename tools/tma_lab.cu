// tma_lab.cu — isolates the TMA box load used by stream.cuh: a [rows][row_bytes] byte matrix described as UINT32 / UINT64
// elements, one box of 32 rows x pitch bytes at an arbitrary (element-aligned) inner offset, tensor map in global
// memory or as a __grid_constant__ parameter.  Checks the bytes that land in shared memory (zero fill past the edges).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool FENCE>
__device__ void run(const void* tmap, int c0, int c1, int bytes, uint8_t* out, uint8_t* smem, unsigned long long* bar) {
    const uint32_t b = smem_u32(bar), dst = smem_u32(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32) {
        if (FENCE) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(tmap), "r"(c0), "r"(c1), "r"(b)
                     : "memory");
    }
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000LL) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = ok ? smem[i] : 0xEE;
}
__global__ void k_global(const void* tmap, int c0, int c1, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    run<true>(tmap, c0, c1, bytes, out, smem, &bar);
}
__global__ void k_param(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    run<false>(&tmap, c0, c1, bytes, out, smem, &bar);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int one(EncodeTiledFn encode, const char* what, int rows, int row_bytes, int es, int pitch, int off_bytes, int row0, bool param) {
    std::vector<uint8_t> h((size_t)rows * row_bytes);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 131 + (i >> 8) * 7 + 1);
    uint8_t *d, *out;
    CK(cudaMalloc(&d, h.size() + 256));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    const int bytes = pitch * 32;
    CK(cudaMalloc(&out, bytes));
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / es), (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(pitch / es), 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-28s encode failed %d\n", what, (int)r); return 0; }
    CUtensorMap* dtm;
    CK(cudaMalloc(&dtm, sizeof(tm)));
    CK(cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    if (param) {
        CK(cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        k_param<<<1, 128, bytes>>>(tm, off_bytes / es, row0, bytes, out);
    } else {
        CK(cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        k_global<<<1, 128, bytes>>>(dtm, off_bytes / es, row0, bytes, out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s %s: %s\n", what, param ? "param " : "global", cudaGetErrorString(e)); return 2; }
    std::vector<uint8_t> got(bytes);
    CK(cudaMemcpy(got.data(), out, bytes, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int rr = 0; rr < 32; rr++)
        for (int b = 0; b < pitch; b++) {
            const int row = row0 + rr, col = off_bytes + b;
            const uint8_t want = (row < rows && col < row_bytes) ? h[(size_t)row * row_bytes + col] : 0;
            if (got[rr * pitch + b] != want) bad++;
        }
    printf("%-28s %s: %ld bad bytes of %d%s\n", what, param ? "param " : "global", bad, bytes, got[0] == 0xEE && bad ? " (timed out)" : "");
    cudaFree(d); cudaFree(out); cudaFree(dtm);
    return 0;
}

int main(int argc, char** argv) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int which = argc > 1 ? atoi(argv[1]) : 0, param = argc > 2 ? atoi(argv[2]) : 1;
    switch (which) {
        case 0: return one(enc, "q4k 576 off 0", 100, 2304, 4, 576, 0, 0, param);
        case 1: return one(enc, "q4k 576 off 1152 row 96", 100, 2304, 4, 576, 1152, 96, param);
        case 2: return one(enc, "q6k 864 off 832", 64, 3360, 4, 864, 832, 32, param);
        case 3: return one(enc, "q6k 864 off 2512 (edge)", 64, 3360, 4, 864, 2512, 32, param);
        case 4: return one(enc, "q80 816 u64 off 816", 64, 4352, 8, 816, 816, 0, param);
        case 5: return one(enc, "q5k 704 off 704", 64, 2816, 4, 704, 704, 0, param);
        case 6: return one(enc, "q6k 848 off 1680", 64, 3360, 4, 848, 1680, 32, param);
        case 7: return one(enc, "q6k 848 off 840 (unaligned)", 64, 3360, 4, 848, 840, 32, param);
    }
    return 0;
}
