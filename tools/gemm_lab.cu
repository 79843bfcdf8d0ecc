// gemm_lab — test / measurement harness for the two tcgen05 dequant-GEMM kernels of the product library
// (csrc/gemm_umma.cuh: first kernel; csrc/gemm_umma2.cuh: warp-specialised persistent kernel).  Not part of the product.
//
//   gemm_lab <type 12|13|14|8> <n_rows> <K> <T> [iters] [ng]
//
// Random GGUF blocks (finite f16 scales) and fp16 activations; the new kernel is compared with the first kernel (itself
// parity-tested against the oracle, tests/test_gpu_ops.py) and both are timed with CUDA events.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../llama-gguf_b200/csrc/common.cuh"
#include "../llama-gguf_b200/csrc/quant.cuh"
#include "../llama-gguf_b200/csrc/gemm_umma.cuh"
#include "../llama-gguf_b200/csrc/gemm_umma2.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static uint16_t f2h_host(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }

static int lab_pitch(int type) {   // stream_pitch(type, 1) of csrc/stream.cuh
    const int raw = 256 / type_block_elems(type) * type_block_bytes(type);
    int maxres = 0;
    for (int ce = 0; ce < 16; ce++) maxres = std::max(maxres, (ce * raw) & 15);
    return (raw + maxres + 15) & ~15;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: gemm_lab type n_rows K T [iters] [ng] [dbg mask: 1 no dequant, 2 no raw loads, 4 no x loads, 8 no MMA]\n"); return 2; }
    const int type = atoi(argv[1]), n_rows = atoi(argv[2]), K = atoi(argv[3]), T = atoi(argv[4]);
    const int iters = argc > 5 ? atoi(argv[5]) : 20, ng = argc > 6 ? atoi(argv[6]) : 2, dbg = argc > 7 ? atoi(argv[7]) : 0;
    const int be = type_block_elems(type), bb = type_block_bytes(type);
    const long long row_bytes = (long long)K / be * bb;
    std::vector<uint8_t> w((size_t)n_rows * row_bytes + 256);
    for (auto& b : w) b = (uint8_t)rnd();
    for (int j = 0; j < n_rows; j++)
        for (int b = 0; b < K / be; b++) {   // finite scales: d ~ 0.01, dmin ~ 0.005
            uint8_t* blk = w.data() + (size_t)j * row_bytes + (size_t)b * bb;
            const uint16_t d = f2h_host(0.002f + 0.00001f * (float)(rnd() % 1000)), dm = f2h_host(0.001f + 0.00001f * (float)(rnd() % 500));
            if (type == T_Q4_K || type == T_Q5_K) { memcpy(blk, &d, 2); memcpy(blk + 2, &dm, 2); }
            else if (type == T_Q6_K) memcpy(blk + 208, &d, 2);
            else memcpy(blk, &d, 2);
        }
    std::vector<uint16_t> x((size_t)T * K);
    for (auto& v : x) v = f2h_host(((float)(rnd() % 2001) - 1000.0f) / 1000.0f);
    uint8_t* dw; __half* dx; float *y0, *y1, *scratch; int* derr;
    CK(cudaMalloc(&dw, w.size())); CK(cudaMalloc(&dx, x.size() * 2));
    CK(cudaMalloc(&y0, (size_t)T * n_rows * 4)); CK(cudaMalloc(&y1, (size_t)T * n_rows * 4));
    const size_t scratch_floats = (size_t)8 * 64 * n_rows;
    CK(cudaMalloc(&scratch, scratch_floats * 4)); CK(cudaMalloc(&derr, 32)); CK(cudaMemset(derr, 0, 32));
    CK(cudaMemcpy(dw, w.data(), w.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
    // a second, larger copy of the weights so that timed iterations do not hit L2 (rotate over copies)
    const int copies = (int)std::max<size_t>(1, std::min<size_t>(16, ((size_t)300 << 20) / w.size() + 1));
    uint8_t* dwc; CK(cudaMalloc(&dwc, w.size() * copies));
    for (int c = 0; c < copies; c++) CK(cudaMemcpy(dwc + (size_t)c * w.size(), w.data(), w.size(), cudaMemcpyHostToDevice));

    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    Umma2EncodeFn encode = (Umma2EncodeFn)fn;
    const int pitch = lab_pitch(type);
    std::vector<CUtensorMap> hmaps(copies + 1);
    for (int c = 0; c <= copies; c++) {
        const cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 4), (cuuint64_t)n_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)(pitch / 4), (cuuint32_t)kUmmaM}, estr[2] = {1, 1};
        void* base = c == copies ? (void*)dw : (void*)(dwc + (size_t)c * w.size());
        if (encode(&hmaps[c], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            fprintf(stderr, "raw tensor map failed\n"); return 1;
        }
    }
    CUtensorMap* dmaps; CK(cudaMalloc(&dmaps, hmaps.size() * sizeof(CUtensorMap)));
    CK(cudaMemcpy(dmaps, hmaps.data(), hmaps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);

    int* dcnt; CK(cudaMalloc(&dcnt, 8192 * 4)); CK(cudaMemset(dcnt, 0, 8192 * 4));
    auto params = [&](const uint8_t* wp, const CUtensorMap* tm, float* y, bool tma_old) {
        UmmaParams p{};
        p.w = wp; p.row_bytes = row_bytes; p.type = type; p.n_rows = n_rows; p.K = K; p.x = dx; p.ldx = K; p.T = T; p.y = y; p.ldy = n_rows;
        p.err = derr;
        if (tm && (tma_old ? T <= 64 : true)) { p.tmap = tm; p.raw_pitch = pitch; p.raw_bytes = 256 / be * bb; }
        umma_plan_split(p, scratch, scratch_floats, n_sm);
        if (!tma_old && p.k_split && !(dbg & 32)) p.tile_cnt = dcnt;   // persistent kernel: in-kernel split-K reduce (dbg 32: separate kernel)
        return p;
    };
    // ---- correctness: first kernel vs persistent kernel on the same inputs
    {
        UmmaParams p0 = params(dw, dmaps + copies, y0, true);
        CK(umma_launch(p0, 0));
        CK(cudaDeviceSynchronize());
        UmmaParams p1 = params(dw, dmaps + copies, y1, false);
        if (!umma2_eligible(p1)) { fprintf(stderr, "not eligible for the persistent kernel\n"); return 1; }
        CK(umma2_launch(encode, p1, n_sm, 232448 - 2048, 0, ng, dbg));
        CK(cudaDeviceSynchronize());
        int herr[8]; CK(cudaMemcpy(herr, derr, 32, cudaMemcpyDeviceToHost));
        std::vector<float> a((size_t)T * n_rows), b((size_t)T * n_rows);
        CK(cudaMemcpy(a.data(), y0, a.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), y1, b.size() * 4, cudaMemcpyDeviceToHost));
        double mx = 0, md = 0; size_t worst = 0; int bad = 0;
        for (size_t i = 0; i < a.size(); i++) {
            if (!std::isfinite(b[i])) bad++;
            mx = std::max(mx, (double)fabsf(a[i]));
            const double dd = fabs((double)a[i] - (double)b[i]);
            if (dd > md) { md = dd; worst = i; }
        }
        printf("check type %d rows %d K %d T %d split %d: max|y| %.4f max|diff| %.3e rel %.3e (worst t %zu j %zu: %.6f vs %.6f) nonfinite %d err %d\n", type, n_rows,
               K, T, p1.k_split, mx, md, md / (mx + 1e-30), worst / n_rows, worst % n_rows, a[worst], b[worst], bad, herr[0]);
        if (!dbg && (herr[0] || bad || md / (mx + 1e-30) > 2e-4)) { printf("FAIL\n"); return 1; }
    }
    // ---- timing (weights rotate over `copies` so every launch streams them from HBM)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; which++) {
        for (int i = 0; i < 3; i++) {
            UmmaParams p = params(dwc + (size_t)(i % copies) * w.size(), dmaps + (i % copies), y1, which == 0);
            CK(which == 0 ? umma_launch(p, 0) : umma2_launch(encode, p, n_sm, 232448 - 2048, 0, ng, dbg));
        }
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < iters; i++) {
            UmmaParams p = params(dwc + (size_t)(i % copies) * w.size(), dmaps + (i % copies), y1, which == 0);
            CK(which == 0 ? umma_launch(p, 0) : umma2_launch(encode, p, n_sm, 232448 - 2048, 0, ng, dbg));
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1000.0 / iters;
        if (which == 1) { Umma2Plan pl{}; size_t sm = 0; CUtensorMap tmp; UmmaParams p = params(dwc, dmaps, y1, false); umma2_plan(p, 232448 - 2048, pl, sm); umma2_encode_xmap(encode, &tmp, p, pl.x3d); printf("  plan: items %d x %d x %d, stages A %d B %d raw %d, smem %zu, x3d %d\n", pl.n_mt, pl.n_z, pl.n_nt, pl.sa, pl.sb, pl.sr, sm, pl.x3d); }
        printf("%s: %.1f us  %.1f TFLOP/s  %.0f GB/s (weights)\n", which == 0 ? "umma1" : "umma2", us, 2.0 * n_rows * K * T / us * 1e-6,
               (double)n_rows * row_bytes / us * 1e-3);
    }
    {   // one more launch with the role clocks recorded
        long long* dprof; CK(cudaMalloc(&dprof, (size_t)n_sm * 16 * 8)); CK(cudaMemset(dprof, 0, (size_t)n_sm * 16 * 8));
        UmmaParams p = params(dwc, dmaps, y1, false);
        CK(umma2_launch(encode, p, n_sm, 232448 - 2048, 0, ng, dbg, dprof));
        CK(cudaDeviceSynchronize());
        std::vector<long long> hp((size_t)n_sm * 16);
        CK(cudaMemcpy(hp.data(), dprof, hp.size() * 8, cudaMemcpyDeviceToHost));
        const char* names[6] = {"producer", "mma", "dequant0", "dequant1", "dequant2", "epilogue"};
        for (int role = 0; role < 6; role++) {
            double tot = 0, wt = 0, mxt = 0; int nn = 0;
            for (int b = 0; b < n_sm; b++) {
                const long long a = hp[(size_t)(b * 8 + role) * 2], c = hp[(size_t)(b * 8 + role) * 2 + 1];
                if (a) { tot += a; wt += c; mxt = std::max(mxt, (double)a); nn++; }
            }
            if (nn) printf("  role %-9s: mean %.0f clk in loop (max %.0f), %.0f waiting (%.0f %%)\n", names[role], tot / nn, mxt, wt / nn, 100.0 * wt / tot);
        }
    }
    int herr[8]; CK(cudaMemcpy(herr, derr, 32, cudaMemcpyDeviceToHost));
    if (!dbg && herr[0]) { printf("FAIL: err %d after timing\n", herr[0]); return 1; }
    printf("OK\n");
    return 0;
}
