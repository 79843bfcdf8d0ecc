// gemv_lab.cu — standalone GPU lab for the dequant-GEMV kernels (no torch, no python):
//   check : one-hot sweeps (which element mapping is wrong, if any), random-shape parity against a
//           double-precision host dequant-then-dot, fused epilogues, determinism
//   time  : per-shape GB/s with weights rotated through > L2 of replicas, sweeping warps/stages/chunk,
//           next to the CUDA-core V1 kernel; a PDL-chained "layer" sequence
// Build: make -C tools      Run (GPU box): tools/gemv_lab [check|time|all]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../llama-gguf_b200/csrc/gemv.cuh"
#include "../llama-gguf_b200/csrc/gemv_mma.cuh"

using namespace b200;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

// ------------------------------------------------------------------ host helpers
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() {
    g_rng ^= g_rng << 13;
    g_rng ^= g_rng >> 7;
    g_rng ^= g_rng << 17;
    return g_rng;
}
static inline float urand() { return (float)((rnd() >> 40) * (1.0 / 16777216.0)); }
static inline float nrand() {
    float u1 = std::max(urand(), 1e-7f), u2 = urand();
    return sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
}
static float h2f(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, o;
    if (e == 0) {
        if (m == 0) o = s << 31;
        else {
            int sh = 0;
            while (!(m & 1024)) { m <<= 1; sh++; }
            m &= 1023;
            o = (s << 31) | ((uint32_t)(113 - sh) << 23) | (m << 13);
        }
    } else if (e == 31) o = (s << 31) | 0x7F800000u | (m << 13);
    else o = (s << 31) | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &o, 4);
    return f;
}
static uint16_t f2h(float f) {  // round-to-nearest, finite normal range is enough here
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000u;
    int e = (int)((x >> 23) & 255) - 127 + 15;
    uint32_t m = x & 0x7FFFFFu;
    if (e <= 0) return (uint16_t)s;
    if (e >= 31) return (uint16_t)(s | 0x7BFF);
    uint32_t r = m >> 13;
    if ((m & 0x1FFF) > 0x1000 || ((m & 0x1FFF) == 0x1000 && (r & 1))) r++;
    return (uint16_t)(s | ((uint32_t)e << 10)) + (uint16_t)r;
}

static void fill_blocks(int type, uint8_t* dst, size_t nblocks) {
    const int bb = type_block_bytes(type);
    size_t nbytes = nblocks * bb;
    size_t i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t r = rnd();
        memcpy(dst + i, &r, 8);
    }
    for (; i < nbytes; i++) dst[i] = (uint8_t)rnd();
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t* p = dst + b * bb;
        uint16_t d = f2h((0.002f + 0.02f * urand()) * ((rnd() & 1) ? 1.f : -1.f));
        uint16_t dm = f2h(0.002f + 0.02f * urand());
        if (type == T_Q4_K || type == T_Q5_K) { memcpy(p, &d, 2); memcpy(p + 2, &dm, 2); }
        else if (type == T_Q6_K) memcpy(p + 208, &d, 2);
        else if (type == T_Q8_0) memcpy(p, &d, 2);
    }
}
static void smk4(const uint8_t* s, int j, int& sc, int& mn) {
    if (j < 4) { sc = s[j] & 63; mn = s[j + 4] & 63; }
    else { sc = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4); mn = ((s[j + 4] >> 4) & 0xF) | ((s[j] >> 6) << 4); }
}
// dequantise one block to doubles (the reference's formulas, dequant.rs:103-109, 205-356)
static void deq_block(int type, const uint8_t* b, double* out) {
    if (type == T_Q8_0) {
        double d = h2f(*(const uint16_t*)b);
        for (int i = 0; i < 32; i++) out[i] = d * (double)(int8_t)b[2 + i];
    } else if (type == T_Q4_K || type == T_Q5_K) {
        double d = h2f(*(const uint16_t*)b), dm = h2f(*(const uint16_t*)(b + 2));
        const uint8_t* qs = b + (type == T_Q5_K ? 48 : 16);
        for (int j = 0; j < 8; j++) {
            int sc, mn;
            smk4(b + 4, j, sc, mn);
            for (int l = 0; l < 32; l++) {
                uint8_t by = qs[(j >> 1) * 32 + l];
                int q = (j & 1) ? (by >> 4) : (by & 15);
                if (type == T_Q5_K && ((b[16 + l] >> j) & 1)) q += 16;
                out[j * 32 + l] = d * sc * q - dm * mn;
            }
        }
    } else if (type == T_Q6_K) {
        double d = h2f(*(const uint16_t*)(b + 208));
        const int8_t* sc = (const int8_t*)(b + 192);
        for (int e = 0; e < 256; e++) {
            int n = e >> 7, r = e & 127, c = r >> 5, l = r & 31;
            uint8_t lb = b[n * 64 + l + ((c & 1) ? 32 : 0)];
            int nib = (c & 2) ? (lb >> 4) : (lb & 15);
            int hb = (b[128 + n * 32 + l] >> (2 * c)) & 3;
            out[e] = d * sc[n * 8 + (l >> 4) + 2 * c] * ((nib | (hb << 4)) - 32);
        }
    }
}
static double ref_row(int type, const uint8_t* row, int K, const float* x) {
    const int be = type_block_elems(type), bb = type_block_bytes(type);
    double acc = 0, tmp[256];
    for (int b = 0; b < K / be; b++) {
        deq_block(type, row + (size_t)b * bb, tmp);
        for (int i = 0; i < be; i++) acc += tmp[i] * (double)x[b * be + i];
    }
    return acc;
}
static const char* tname(int t) {
    switch (t) { case T_Q4_K: return "Q4_K"; case T_Q5_K: return "Q5_K"; case T_Q6_K: return "Q6_K"; case T_Q8_0: return "Q8_0"; }
    return "?";
}

// ------------------------------------------------------------------ device state
static int g_nsm = 148;
static float* g_part = nullptr;
static unsigned int* g_tickets = nullptr;
static int* g_err = nullptr;
static unsigned long long* g_dbg = nullptr;
static bool g_use_dbg = false;
static size_t g_smem_limit = 227 * 1024;

struct Knobs { int warps = 16, stages = 3; };

static void lab_init() {
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, 0));
    g_nsm = prop.multiProcessorCount;
    g_smem_limit = prop.sharedMemPerBlockOptin - 8192;  // static shared memory of the kernels
    printf("device: %s, %d SMs, smem optin %zu\n", prop.name, g_nsm, (size_t)prop.sharedMemPerBlockOptin);
    CK(cudaMalloc(&g_part, (size_t)g_nsm * kMmaMaxWarps * 2 * 32 * 4));
    CK(cudaMalloc(&g_tickets, 65536 * 4));
    CK(cudaMemset(g_tickets, 0, 65536 * 4));
    CK(cudaMalloc(&g_err, 4));
    CK(cudaMemset(g_err, 0, 4));
    CK(cudaMalloc(&g_dbg, (size_t)g_nsm * kMmaMaxWarps * 8 * 8));
    CK(mma_set_smem_limit((int)g_smem_limit));
    CK(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
}

static bool launch_mma(cudaStream_t st, MParams p, const Knobs& kn, bool pdl) {
    MPlan plan;
    p.part = g_part;
    p.tickets = g_tickets;
    p.err = g_err;
    p.dbg = g_use_dbg ? g_dbg : nullptr;
    if (!mma_plan(p, g_nsm, kn.warps, kn.stages, g_smem_limit, plan)) return false;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(plan.warps * 32);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = pdl ? at : nullptr;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, mma_kernel_for(plan.stages), p));
    return true;
}
static void launch_v1(cudaStream_t st, GemvParams p, bool pdl) {
    int n_tasks;
    if (p.epi == EPI_SWIGLU) n_tasks = (p.seg[0].n_rows + 1) / 2;
    else { n_tasks = 0; for (int s = 0; s < p.n_seg; s++) n_tasks += (p.seg[s].n_rows + kGemvR - 1) / kGemvR; }
    int grid = std::max(1, std::min((n_tasks + kGemvWarps - 1) / kGemvWarps, 2 * g_nsm));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemvThreads);
    cfg.dynamicSmemBytes = (size_t)xpad_floats(p.K) * 4;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = pdl ? at : nullptr;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, gemv_kernel, p));
}
static int read_err() {
    int e = 0;
    CK(cudaMemcpy(&e, g_err, 4, cudaMemcpyDeviceToHost));
    return e;
}

static MSeg mseg(const uint8_t* w, float* out, const float* bias, int type, int K, int n_rows) {
    MSeg s{};
    s.w = w; s.out = out; s.bias = bias; s.type = type; s.n_rows = n_rows;
    s.row_bytes = (long long)(K / type_block_elems(type)) * type_block_bytes(type);
    return s;
}
static GemvSeg vseg(const uint8_t* w, float* out, const float* bias, int type, int K, int n_rows) {
    GemvSeg s{};
    s.w = w; s.out = out; s.bias = bias; s.type = type; s.n_rows = n_rows;
    s.row_bytes = (long long)(K / type_block_elems(type)) * type_block_bytes(type);
    return s;
}

// ------------------------------------------------------------------ checks
static int g_fail = 0;

// y = W e_k for every k: compares against the dequantised element; prints the first few mismatches
static void check_onehot(int type, int K, const Knobs& kn) {
    const int N = 16, be = type_block_elems(type), bb = type_block_bytes(type);
    const size_t rb = (size_t)(K / be) * bb;
    std::vector<uint8_t> hw(rb * N);
    fill_blocks(type, hw.data(), (size_t)N * (K / be));
    uint8_t* dw; float *dx, *dy;
    CK(cudaMalloc(&dw, rb * N + 256)); CK(cudaMemcpy(dw, hw.data(), rb * N, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dx, (size_t)K * K * 4)); CK(cudaMalloc(&dy, (size_t)K * N * 4));
    std::vector<float> hx((size_t)K * K, 0.f);
    for (int k = 0; k < K; k++) hx[(size_t)k * K + k] = 1.0f;
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    for (int k = 0; k < K; k++) {
        MParams p{};
        p.seg[0] = mseg(dw, dy + (size_t)k * N, nullptr, type, K, N);
        p.n_seg = 1; p.K = K; p.x = dx + (size_t)k * K; p.epi = ME_STORE;
        if (!launch_mma(0, p, kn, false)) { printf("onehot %s: plan failed\n", tname(type)); g_fail++; return; }
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> hy((size_t)K * N);
    CK(cudaMemcpy(hy.data(), dy, hy.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double tmp[256];
    for (int j = 0; j < N; j++)
        for (int b = 0; b < K / be; b++) {
            deq_block(type, hw.data() + j * rb + (size_t)b * bb, tmp);
            for (int i = 0; i < be; i++) {
                int k = b * be + i;
                float got = hy[(size_t)k * N + j];
                double want = tmp[i];
                if (fabs(got - want) > 2e-4 * std::max(1.0, fabs(want)) + 1e-6) {
                    if (bad < 12) printf("  onehot %s K=%d: row %d elem %d (blk %d, i %d): got %.6f want %.6f\n", tname(type), K, j, k, b, i, got, want);
                    bad++;
                }
            }
        }
    printf("onehot %-5s K=%-5d w=%d st=%d: %s (%d / %d mismatches) err=%d\n", tname(type), K, kn.warps, kn.stages, bad ? "FAIL" : "ok", bad, K * N, read_err());
    if (bad) g_fail++;
    cudaFree(dw); cudaFree(dx); cudaFree(dy);
}

struct Shape { int type, K, N; };

static double check_random(int type, int K, int N, const Knobs& kn, int epi, bool with_bias, bool with_norm, float xscale = 1.0f) {
    const int be = type_block_elems(type), bb = type_block_bytes(type);
    const size_t rb = (size_t)(K / be) * bb;
    const int nmat = (epi == ME_SWIGLU) ? 2 : 1;
    std::vector<uint8_t> hw(rb * N * nmat);
    fill_blocks(type, hw.data(), (size_t)N * nmat * (K / be));
    std::vector<float> hx(K), hn(K), hb(N), hr(N);
    for (auto& v : hx) v = nrand() * xscale;
    for (auto& v : hn) v = 1.0f + 0.1f * nrand();
    for (auto& v : hb) v = 0.1f * nrand();
    for (auto& v : hr) v = nrand();
    uint8_t* dw; float *dx, *dn, *db, *dr, *dy;
    CK(cudaMalloc(&dw, hw.size() + 256)); CK(cudaMemcpy(dw, hw.data(), hw.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dx, K * 4)); CK(cudaMemcpy(dx, hx.data(), K * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dn, K * 4)); CK(cudaMemcpy(dn, hn.data(), K * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&db, N * 4)); CK(cudaMemcpy(db, hb.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dr, N * 4)); CK(cudaMemcpy(dr, hr.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dy, N * 4)); CK(cudaMemset(dy, 0xFF, N * 4));
    MParams p{};
    p.seg[0] = mseg(dw, dy, with_bias ? db : nullptr, type, K, N);
    p.n_seg = 1;
    if (epi == ME_SWIGLU) { p.seg[1] = mseg(dw + rb * N, dy, nullptr, type, K, N); p.n_seg = 2; }
    p.K = K; p.x = dx; p.norm_w = with_norm ? dn : nullptr; p.eps = 1e-5f; p.epi = epi; p.residual = dr;
    std::vector<float> y1(N), y2(N);
    bool okp = launch_mma(0, p, kn, false);
    CK(cudaDeviceSynchronize());
    if (!okp) { printf("random %s K=%d N=%d: plan failed\n", tname(type), K, N); g_fail++; return -1; }
    CK(cudaMemcpy(y1.data(), dy, N * 4, cudaMemcpyDeviceToHost));
    launch_mma(0, p, kn, false);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(y2.data(), dy, N * 4, cudaMemcpyDeviceToHost));
    const bool det = memcmp(y1.data(), y2.data(), N * 4) == 0;
    // reference
    std::vector<float> xe(hx);
    if (with_norm) {
        double ss = 0;
        for (int i = 0; i < K; i++) ss += (double)hx[i] * hx[i];
        float inv = 1.0f / sqrtf((float)(ss / K) + 1e-5f);
        for (int i = 0; i < K; i++) xe[i] = (hx[i] * inv) * hn[i];
    }
    const int nsample = std::min(N, 256);
    double max_err = 0, rms = 0;
    std::vector<double> want(nsample);
    std::vector<int> rows(nsample);
    for (int i = 0; i < nsample; i++) {
        int j = (nsample == N) ? i : (int)(rnd() % N);
        if (i == 0) j = 0;
        if (i == 1) j = N - 1;
        rows[i] = j;
        double v = ref_row(type, hw.data() + (size_t)j * rb, K, xe.data());
        if (epi == ME_SWIGLU) {
            double u = ref_row(type, hw.data() + (size_t)(N + j) * rb, K, xe.data());
            v = v / (1.0 + exp(-v)) * u;
        }
        if (with_bias) v += hb[j];
        if (epi == ME_RESIDUAL) v += hr[j];
        want[i] = v;
        rms += v * v;
    }
    rms = sqrt(rms / nsample);
    int nanc = 0;
    for (int i = 0; i < nsample; i++) {
        float got = y1[rows[i]];
        if (got != got) nanc++;
        max_err = std::max(max_err, fabs(got - want[i]) / (rms + 1e-30));
    }
    const bool ok = max_err < 2e-4 && det && nanc == 0 && read_err() == 0;
    printf("random %-5s K=%-6d N=%-7d epi=%d bias=%d norm=%d xs=%g w=%d st=%d: max|err|/rms=%.2e det=%d nan=%d err=%d %s\n", tname(type), K, N, epi,
           with_bias, with_norm, xscale, kn.warps, kn.stages, max_err, det, nanc, read_err(), ok ? "ok" : "FAIL");
    if (!ok) g_fail++;
    cudaFree(dw); cudaFree(dx); cudaFree(dn); cudaFree(db); cudaFree(dr); cudaFree(dy);
    return max_err;
}

// three segments of different types sharing x (the QKV launch)
static void check_qkv(const Knobs& kn) {
    const int K = 4096, Ns[3] = {4096, 1024, 1000};
    const int types[3] = {T_Q4_K, T_Q4_K, T_Q6_K};
    std::vector<uint8_t> hw[3];
    uint8_t* dw[3]; float* dy[3];
    std::vector<float> hx(K);
    for (auto& v : hx) v = nrand();
    float* dx; CK(cudaMalloc(&dx, K * 4)); CK(cudaMemcpy(dx, hx.data(), K * 4, cudaMemcpyHostToDevice));
    MParams p{};
    for (int s = 0; s < 3; s++) {
        const int be = type_block_elems(types[s]), bb = type_block_bytes(types[s]);
        hw[s].resize((size_t)Ns[s] * (K / be) * bb);
        fill_blocks(types[s], hw[s].data(), (size_t)Ns[s] * (K / be));
        CK(cudaMalloc(&dw[s], hw[s].size() + 256)); CK(cudaMemcpy(dw[s], hw[s].data(), hw[s].size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dy[s], Ns[s] * 4)); CK(cudaMemset(dy[s], 0xFF, Ns[s] * 4));
        p.seg[s] = mseg(dw[s], dy[s], nullptr, types[s], K, Ns[s]);
    }
    p.n_seg = 3; p.K = K; p.x = dx; p.epi = ME_STORE;
    launch_mma(0, p, kn, false);
    CK(cudaDeviceSynchronize());
    double worst = 0;
    for (int s = 0; s < 3; s++) {
        std::vector<float> y(Ns[s]);
        CK(cudaMemcpy(y.data(), dy[s], Ns[s] * 4, cudaMemcpyDeviceToHost));
        const size_t rb = hw[s].size() / Ns[s];
        double rms = 0, me = 0;
        std::vector<double> want(Ns[s]);
        for (int j = 0; j < Ns[s]; j += 7) { want[j] = ref_row(types[s], hw[s].data() + j * rb, K, hx.data()); rms += want[j] * want[j]; }
        rms = sqrt(rms / ((Ns[s] + 6) / 7));
        for (int j = 0; j < Ns[s]; j += 7) { double e = fabs(y[j] - want[j]) / rms; if (!(e <= me)) me = e; }
        worst = std::max(worst, me);
        cudaFree(dw[s]); cudaFree(dy[s]);
    }
    const bool ok = worst < 2e-4 && read_err() == 0;
    printf("qkv 3-seg (Q4_K,Q4_K,Q6_K; last N=1000): max|err|/rms=%.2e err=%d %s\n", worst, read_err(), ok ? "ok" : "FAIL");
    if (!ok) g_fail++;
    cudaFree(dx);
}

// ------------------------------------------------------------------ timing
struct TimeCase { const char* name; int type; int K; int N; int epi; };

static void time_case(const TimeCase& tc, const std::vector<Knobs>& knobs, bool also_v1) {
    const int be = type_block_elems(tc.type), bb = type_block_bytes(tc.type);
    const size_t rb = (size_t)(tc.K / be) * bb;
    const int nmat = tc.epi == ME_SWIGLU ? 2 : 1;
    const size_t wbytes = rb * tc.N * nmat;
    const int reps = (int)std::max<size_t>(2, (size_t)600e6 / wbytes + 1);
    uint8_t* dw; float *dx, *dy, *dr;
    CK(cudaMalloc(&dw, wbytes * reps + 256));
    {   // one random replica on the host, copied `reps` times
        std::vector<uint8_t> hw(wbytes);
        fill_blocks(tc.type, hw.data(), wbytes / bb);
        for (int r = 0; r < reps; r++) CK(cudaMemcpy(dw + wbytes * r, hw.data(), wbytes, cudaMemcpyHostToDevice));
    }
    std::vector<float> hx(tc.K);
    for (auto& v : hx) v = nrand();
    CK(cudaMalloc(&dx, tc.K * 4)); CK(cudaMemcpy(dx, hx.data(), tc.K * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dy, (size_t)tc.N * 4)); CK(cudaMalloc(&dr, (size_t)tc.N * 4)); CK(cudaMemset(dr, 0, (size_t)tc.N * 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = std::max(reps * 2, 20);
    auto run = [&](auto&& launch_one) {
        for (int i = 0; i < reps; i++) launch_one(i % reps);  // warm-up pass
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; i++) launch_one(i % reps);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        return ms / iters;
    };
    printf("%-26s %-5s K=%-6d N=%-7d %7.2f MB x%d reps\n", tc.name, tname(tc.type), tc.K, tc.N, wbytes / 1e6, reps);
    for (const Knobs& kn : knobs) {
        for (int pdl = 0; pdl < 2; pdl++) {
            bool okp = true;
            float ms = run([&](int r) {
                MParams p{};
                p.seg[0] = mseg(dw + wbytes * r, dy, nullptr, tc.type, tc.K, tc.N);
                p.n_seg = 1;
                if (tc.epi == ME_SWIGLU) { p.seg[1] = mseg(dw + wbytes * r + rb * tc.N, dy, nullptr, tc.type, tc.K, tc.N); p.n_seg = 2; }
                p.K = tc.K; p.x = dx; p.epi = tc.epi; p.residual = dr;
                okp = launch_mma(st, p, kn, pdl != 0) && okp;
            });
            MParams pp{}; MPlan plan{};
            pp.seg[0] = mseg(dw, dy, nullptr, tc.type, tc.K, tc.N); pp.n_seg = 1;
            if (tc.epi == ME_SWIGLU) { pp.seg[1] = pp.seg[0]; pp.n_seg = 2; }
            pp.K = tc.K; pp.epi = tc.epi;
            mma_plan(pp, g_nsm, kn.warps, kn.stages, g_smem_limit, plan);
            printf("   mma w=%-2d st=%d (got w=%d st=%d smem=%zuK grid=%d) pdl=%d: %8.2f us  %7.1f GB/s%s err=%d\n", kn.warps, kn.stages,
                   plan.warps, plan.stages, plan.smem / 1024, plan.grid, pdl, ms * 1e3, wbytes / (ms * 1e-3) / 1e9, okp ? "" : " PLAN-FAILED", read_err());
        }
    }
    if (also_v1) {
        float ms = run([&](int r) {
            GemvParams p{};
            p.seg[0] = vseg(dw + wbytes * r, dy, nullptr, tc.type, tc.K, tc.N);
            p.n_seg = 1;
            if (tc.epi == ME_SWIGLU) { p.seg[1] = vseg(dw + wbytes * r + rb * tc.N, dy, nullptr, tc.type, tc.K, tc.N); p.n_seg = 2; }
            p.K = tc.K; p.x = dx; p.epi = tc.epi == ME_SWIGLU ? EPI_SWIGLU : (tc.epi == ME_RESIDUAL ? EPI_RESIDUAL : EPI_STORE); p.residual = dr;
            launch_v1(st, p, true);
        });
        printf("   v1 cuda-core pdl=1: %8.2f us  %7.1f GB/s\n", ms * 1e3, wbytes / (ms * 1e-3) / 1e9);
    }
    cudaFree(dw); cudaFree(dx); cudaFree(dy); cudaFree(dr);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
}


// in-kernel globaltimer stamps of one launch: where does the time of a kernel go?
static void timeline_case(const TimeCase& tc, const Knobs& kn) {
    const int be = type_block_elems(tc.type), bb = type_block_bytes(tc.type);
    const size_t rb = (size_t)(tc.K / be) * bb;
    const int nmat = tc.epi == ME_SWIGLU ? 2 : 1;
    const size_t wbytes = rb * tc.N * nmat;
    uint8_t* dw; float *dx, *dy, *dr; void* flush;
    CK(cudaMalloc(&dw, wbytes + 256)); CK(cudaMemset(dw, 0x11, wbytes));
    CK(cudaMalloc(&dx, tc.K * 4)); CK(cudaMemset(dx, 0, tc.K * 4));
    CK(cudaMalloc(&dy, (size_t)tc.N * 4)); CK(cudaMalloc(&dr, (size_t)tc.N * 4)); CK(cudaMemset(dr, 0, (size_t)tc.N * 4));
    CK(cudaMalloc(&flush, 512u << 20));
    g_use_dbg = true;
    const size_t nw = (size_t)g_nsm * kMmaMaxWarps;
    std::vector<unsigned long long> h(nw * 8);
    const char* names[7] = {"start", "pdl_wait done", "x loaded+copies issued", "x staged", "first unit landed", "last unit computed", "exit"};
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(flush, rep, 512u << 20));
        CK(cudaMemset(g_dbg, 0, nw * 64));
        CK(cudaDeviceSynchronize());
        MParams p{};
        p.seg[0] = mseg(dw, dy, nullptr, tc.type, tc.K, tc.N); p.n_seg = 1;
        if (tc.epi == ME_SWIGLU) { p.seg[1] = mseg(dw + rb * tc.N, dy, nullptr, tc.type, tc.K, tc.N); p.n_seg = 2; }
        p.K = tc.K; p.x = dx; p.epi = tc.epi; p.residual = dr;
        launch_mma(0, p, kn, false);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), g_dbg, nw * 64, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (size_t w = 0; w < nw; w++) if (h[w * 8]) t0 = std::min(t0, h[w * 8]);
        printf("timeline %s w=%d st=%d rep %d (ns after the first warp's start; min / avg / max over warps)\n", tc.name, kn.warps, kn.stages, rep);
        for (int i = 0; i < 7; i++) {
            unsigned long long mn = ~0ull, mx = 0; double sum = 0; int n = 0;
            for (size_t w = 0; w < nw; w++) {
                if (!h[w * 8] || !h[w * 8 + i]) continue;
                unsigned long long d = h[w * 8 + i] - t0;
                mn = std::min(mn, d); mx = std::max(mx, d); sum += d; n++;
            }
            if (n) printf("   %-24s %7llu / %9.0f / %7llu   (%d warps)\n", names[i], mn, sum / n, mx, n);
        }
    }
    g_use_dbg = false;
    cudaFree(dw); cudaFree(dx); cudaFree(dy); cudaFree(dr); cudaFree(flush);
}

// a Llama-3-8B layer's four GEMVs back to back with PDL, over `layers` distinct weight sets
static void time_layer_chain(const Knobs& kn, int down_type, bool graph) {
    const int H = 4096, I = 14336, layers = 8;
    struct W { uint8_t *q, *k, *v, *o, *g, *u, *d; };
    auto bytes = [](int type, int K, int N) { return (size_t)(K / type_block_elems(type)) * type_block_bytes(type) * N; };
    std::vector<W> ws(layers);
    size_t per_layer = bytes(T_Q4_K, H, 4096) * 2 + bytes(T_Q4_K, H, 1024) + bytes(down_type, H, 1024) + bytes(T_Q4_K, H, I) * 2 + bytes(down_type, I, H);
    auto mk = [&](int type, int K, int N) {
        size_t b = bytes(type, K, N);
        std::vector<uint8_t> h(b);
        fill_blocks(type, h.data(), b / type_block_bytes(type));
        uint8_t* d; CK(cudaMalloc(&d, b + 256)); CK(cudaMemcpy(d, h.data(), b, cudaMemcpyHostToDevice));
        return d;
    };
    // one random set on the host per tensor kind, replicated per layer by device copies
    W proto{mk(T_Q4_K, H, 4096), mk(T_Q4_K, H, 1024), mk(down_type, H, 1024), mk(T_Q4_K, H, 4096), mk(T_Q4_K, H, I), mk(T_Q4_K, H, I), mk(down_type, I, H)};
    for (int l = 0; l < layers; l++) {
        auto dup = [&](uint8_t* src, size_t b) { uint8_t* d; CK(cudaMalloc(&d, b + 256)); CK(cudaMemcpy(d, src, b, cudaMemcpyDeviceToDevice)); return d; };
        ws[l] = W{dup(proto.q, bytes(T_Q4_K, H, 4096)), dup(proto.k, bytes(T_Q4_K, H, 1024)), dup(proto.v, bytes(down_type, H, 1024)),
                  dup(proto.o, bytes(T_Q4_K, H, 4096)), dup(proto.g, bytes(T_Q4_K, H, I)), dup(proto.u, bytes(T_Q4_K, H, I)), dup(proto.d, bytes(down_type, I, H))};
    }
    float *xa, *xb, *qkv, *hb, *nw;
    CK(cudaMalloc(&xa, H * 4)); CK(cudaMalloc(&xb, H * 4)); CK(cudaMalloc(&qkv, 6144 * 4)); CK(cudaMalloc(&hb, I * 4)); CK(cudaMalloc(&nw, H * 4));
    std::vector<float> hx(H);
    for (auto& v : hx) v = nrand();
    CK(cudaMemcpy(xa, hx.data(), H * 4, cudaMemcpyHostToDevice));
    for (auto& v : hx) v = 1.0f;
    CK(cudaMemcpy(nw, hx.data(), H * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(hb, 0, I * 4)); CK(cudaMemset(qkv, 0, 6144 * 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    auto enqueue = [&]() {
        for (int l = 0; l < layers; l++) {
            const W& w = ws[l];
            MParams p{};
            p.seg[0] = mseg(w.q, qkv, nullptr, T_Q4_K, H, 4096); p.seg[1] = mseg(w.k, qkv + 4096, nullptr, T_Q4_K, H, 1024);
            p.seg[2] = mseg(w.v, qkv + 5120, nullptr, down_type, H, 1024);
            p.n_seg = 3; p.K = H; p.x = xa; p.norm_w = nw; p.eps = 1e-5f; p.epi = ME_STORE;
            launch_mma(st, p, kn, true);
            MParams o{};
            o.seg[0] = mseg(w.o, xb, nullptr, T_Q4_K, H, 4096); o.n_seg = 1; o.K = H; o.x = qkv; o.epi = ME_RESIDUAL; o.residual = xa;
            launch_mma(st, o, kn, true);
            MParams gu{};
            gu.seg[0] = mseg(w.g, hb, nullptr, T_Q4_K, H, I); gu.seg[1] = mseg(w.u, hb, nullptr, T_Q4_K, H, I);
            gu.n_seg = 2; gu.K = H; gu.x = xb; gu.norm_w = nw; gu.eps = 1e-5f; gu.epi = ME_SWIGLU;
            launch_mma(st, gu, kn, true);
            MParams dn{};
            dn.seg[0] = mseg(w.d, xa, nullptr, down_type, I, H); dn.n_seg = 1; dn.K = I; dn.x = hb; dn.epi = ME_RESIDUAL; dn.residual = xb;
            launch_mma(st, dn, kn, true);
        }
    };
    cudaGraphExec_t gx = nullptr;
    if (graph) {
        cudaGraph_t gph;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        enqueue();
        CK(cudaStreamEndCapture(st, &gph));
        CK(cudaGraphInstantiate(&gx, gph, 0));
        cudaGraphDestroy(gph);
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto once = [&]() { if (graph) CK(cudaGraphLaunch(gx, st)); else enqueue(); };
    once(); once();
    CK(cudaStreamSynchronize(st));
    const int iters = 10;
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; i++) once();
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us_layer = ms * 1e3 / iters / layers;
    printf("layer chain (down/v=%s, %s, w=%d st=%d): %.2f us/layer, %.1f MB/layer -> %.1f GB/s  (x32 layers = %.3f ms) err=%d\n", tname(down_type),
           graph ? "graph+PDL" : "stream+PDL", kn.warps, kn.stages, us_layer, per_layer / 1e6, per_layer / (us_layer * 1e-6) / 1e9, us_layer * 32e-3, read_err());
    if (gx) cudaGraphExecDestroy(gx);
    for (auto& w : ws) { cudaFree(w.q); cudaFree(w.k); cudaFree(w.v); cudaFree(w.o); cudaFree(w.g); cudaFree(w.u); cudaFree(w.d); }
    cudaFree(proto.q); cudaFree(proto.k); cudaFree(proto.v); cudaFree(proto.o); cudaFree(proto.g); cudaFree(proto.u); cudaFree(proto.d);
    cudaFree(xa); cudaFree(xb); cudaFree(qkv); cudaFree(hb); cudaFree(nw);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
}


// ------------------------------------------------------------------ raw streaming probes
// How fast can one SM pull bytes with each mechanism, in the ring structure of the GEMV (per-warp rings,
// `rows` copies of `S` bytes per stage)?  No arithmetic beyond one xor per lane per stage.
struct StreamParams {
    const uint8_t* src;
    long long total;       // bytes
    int S;                 // bytes per copy (multiple of 16)
    int rows;              // copies per stage
    int stages;
    unsigned int* sink;
};

__global__ void __launch_bounds__(512, 1) stream_bulk_kernel(const StreamParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[16 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t stage_bytes = (uint32_t)p.S * p.rows;
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * p.stages * stage_bytes;
    const uint32_t wbar = smem_u32(&bars[warp * 8]);
    if (lane == 0) for (int s = 0; s < p.stages; s++) mbar_init(wbar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const long long W = (long long)gridDim.x * nw, gw = (long long)warp * gridDim.x + blockIdx.x;
    const long long n_stage_total = p.total / stage_bytes;
    const long long k0 = gw * n_stage_total / W, k1 = (gw + 1) * n_stage_total / W;
    const int n = (int)(k1 - k0);
    auto issue = [&](int k, int st) {
        const uint8_t* base = p.src + (k0 + k) * (long long)stage_bytes;
        if (lane == 0) mbar_arrive_expect_tx(wbar + 8 * st, stage_bytes);
        __syncwarp();
        if (lane < p.rows) bulk_g2s(ring + st * stage_bytes + lane * p.S, base + (long long)lane * p.S, p.S, wbar + 8 * st);
    };
    const int pre = min(p.stages - 1, n);
    for (int k = 0; k < pre; k++) issue(k, k);
    unsigned int acc = 0;
    for (int k = 0; k < n; k++) {
        if (k + p.stages - 1 < n) { __syncwarp(); issue(k + p.stages - 1, (k + p.stages - 1) % p.stages); }
        const int st = k % p.stages;
        if (!mbar_wait(wbar + 8 * st, (k / p.stages) & 1, nullptr)) break;
        acc ^= lds32(ring + st * stage_bytes + lane * 4);
    }
    if (acc == 0x12345678u) p.sink[0] = acc;
}

// cp.async (LDGSTS) 16 B per lane, commit groups per stage
__global__ void __launch_bounds__(512, 1) stream_cpasync_kernel(const StreamParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t stage_bytes = (uint32_t)p.S * p.rows;
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * p.stages * stage_bytes;
    const long long W = (long long)gridDim.x * nw, gw = (long long)warp * gridDim.x + blockIdx.x;
    const long long n_stage_total = p.total / stage_bytes;
    const long long k0 = gw * n_stage_total / W, k1 = (gw + 1) * n_stage_total / W;
    const int n = (int)(k1 - k0);
    const int pieces = stage_bytes / 16;
    auto issue = [&](int k, int st) {
        const uint8_t* base = p.src + (k0 + k) * (long long)stage_bytes;
        for (int q = lane; q < pieces; q += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + st * stage_bytes + q * 16), "l"(base + q * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int k = 0; k < p.stages - 1; k++) { if (k < n) issue(k, k); else asm volatile("cp.async.commit_group;" ::: "memory"); }
    unsigned int acc = 0;
    for (int k = 0; k < n; k++) {
        if (k + p.stages - 1 < n) issue(k + p.stages - 1, (k + p.stages - 1) % p.stages);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        if (p.stages == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (p.stages == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else asm volatile("cp.async.wait_group 3;" ::: "memory");
        __syncwarp();
        acc ^= lds32(ring + (k % p.stages) * stage_bytes + lane * 4);
        __syncwarp();
    }
    if (acc == 0x12345678u) p.sink[0] = acc;
}

// plain LDG.128 into registers, `S/16` loads in flight per lane
template <int U>
__global__ void __launch_bounds__(512, 1) stream_ldg_kernel(const StreamParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long W = (long long)gridDim.x * nw, gw = (long long)warp * gridDim.x + blockIdx.x;
    const long long chunk = 512LL * U;
    const long long n_total = p.total / chunk;
    const long long k0 = gw * n_total / W, k1 = (gw + 1) * n_total / W;
    unsigned int acc = 0;
    for (long long k = k0; k < k1; k++) {
        const uint8_t* base = p.src + k * chunk + lane * 16;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ldg_stream_u4(base + u * 512);
#pragma unroll
        for (int u = 0; u < U; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) p.sink[0] = acc;
}


// the GEMV's exact access pattern (16 rows at stride row_bytes, S bytes per row per stage, walking along K, then
// the next 16-row tile), cp.async 16 B pieces, no arithmetic: separates the memory pattern from the compute
__global__ void __launch_bounds__(512, 1) stream_rows_kernel(const uint8_t* src, int n_rows, int row_bytes, int S, int stages, unsigned int* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
    const int RS = (S + 63) & ~63;
    const uint32_t stage_bytes = 16u * RS;
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * stages * stage_bytes;
    const int chunks = row_bytes / S, tiles = n_rows / 16;
    const long long U = (long long)chunks * tiles, W = (long long)gridDim.x * nw, gw = (long long)warp * gridDim.x + blockIdx.x;
    const long long u0 = gw * U / W, u1 = (gw + 1) * U / W;
    const int n = (int)(u1 - u0);
    auto issue = [&](long long u, int st) {
        const int tile = (int)(u / chunks), chunk = (int)(u - (long long)tile * chunks);
        const uint8_t* sa = src + (long long)(tile * 16 + g) * row_bytes + (long long)chunk * S + 16 * t;
        const uint8_t* sb = sa + 8LL * row_bytes;
        const uint32_t dst = ring + st * stage_bytes + g * RS + 16 * t;
        for (int i = 0; i < (S + 63) / 64; i++) {
            cp_async16(dst + 64 * i, sa + 64 * i);
            cp_async16(dst + 8 * RS + 64 * i, sb + 64 * i);
        }
    };
    for (int k = 0; k < stages - 1; k++) { if (k < n) issue(u0 + k, k); cp_async_commit(); }
    unsigned int acc = 0;
    for (int k = 0; k < n; k++) {
        if (k + stages - 1 < n) issue(u0 + k + stages - 1, (k + stages - 1) % stages);
        cp_async_commit();
        if (stages == 2) cp_async_wait<1>(); else if (stages == 3) cp_async_wait<2>(); else cp_async_wait<3>();
        __syncwarp();
        acc ^= lds32(ring + (k % stages) * stage_bytes + lane * 4);
        __syncwarp();
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
static void stream_rows_probe() {
    const int row_bytes = 2304, n_rows = 14336 * 2 * 8;   // 8 gate/up-sized matrices back to back: 528 MB
    const long long total = (long long)row_bytes * n_rows;
    uint8_t* src; unsigned int* sink;
    CK(cudaMalloc(&src, total + 4096)); CK(cudaMemset(src, 1, total)); CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(stream_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem_limit));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("--- GEMV access pattern, copy only (rows of %d B, 16-row tiles), GB/s\n", row_bytes);
    struct Cfg { int S, warps, stages; };
    for (Cfg c : {Cfg{288, 16, 2}, Cfg{288, 12, 3}, Cfg{576, 8, 2}, Cfg{576, 10, 2}, Cfg{1152, 6, 2}, Cfg{2304, 4, 2}, Cfg{2304, 2, 3}}) {
        const int RS = (c.S + 63) & ~63;
        size_t sm = (size_t)16 * RS * c.stages * c.warps;
        if (sm > g_smem_limit) { printf("rows S=%d w=%d st=%d: smem too large\n", c.S, c.warps, c.stages); continue; }
        stream_rows_kernel<<<g_nsm, c.warps * 32, sm>>>(src, n_rows, row_bytes, c.S, c.stages, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; i++) stream_rows_kernel<<<g_nsm, c.warps * 32, sm>>>(src, n_rows, row_bytes, c.S, c.stages, sink);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("rows    S=%-5d w=%-2d st=%d (ring %3zuK/SM): %7.1f\n", c.S, c.warps, c.stages, sm / 1024, total / (ms / 3 * 1e-3) / 1e9);
    }
    cudaFree(src); cudaFree(sink);
}

static void stream_probes() {
    const long long total = 1LL << 30;
    uint8_t* src; unsigned int* sink;
    CK(cudaMalloc(&src, total + 4096)); CK(cudaMemset(src, 1, total)); CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(stream_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem_limit));
    CK(cudaFuncSetAttribute(stream_cpasync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem_limit));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto&& f) {
        f(); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; i++) f();
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        return total / (ms / 3 * 1e-3) / 1e9;
    };
    printf("--- streaming probes over 1 GiB (GB/s); 148 CTAs\n");
    struct Cfg { int S, rows, warps, stages; };
    for (Cfg c : {Cfg{288, 16, 16, 2}, Cfg{576, 16, 8, 2}, Cfg{576, 16, 12, 2}, Cfg{1152, 16, 4, 2}, Cfg{1152, 16, 6, 2}, Cfg{2304, 16, 4, 2}, Cfg{4608, 1, 16, 2},
                  Cfg{4608, 1, 16, 3}, Cfg{9216, 1, 8, 2}, Cfg{9216, 1, 12, 2}, Cfg{18432, 1, 4, 3}, Cfg{2304, 2, 16, 2}, Cfg{1152, 4, 16, 2}, Cfg{576, 8, 16, 2}, Cfg{288, 16, 16, 3},
                  Cfg{288, 16, 8, 4}, Cfg{288, 16, 8, 2}}) {
        StreamParams p{src, total, c.S, c.rows, c.stages, sink};
        size_t sm = (size_t)c.S * c.rows * c.stages * c.warps;
        if (sm > g_smem_limit) { printf("bulk S=%d rows=%d w=%d st=%d: smem too large\n", c.S, c.rows, c.warps, c.stages); continue; }
        double gbs = timeit([&]() { stream_bulk_kernel<<<g_nsm, c.warps * 32, sm>>>(p); });
        printf("bulk    S=%-5d rows=%-2d w=%-2d st=%d (ring %3zuK/SM): %7.1f\n", c.S, c.rows, c.warps, c.stages, sm / 1024, gbs);
    }
    for (Cfg c : {Cfg{288, 16, 16, 2}, Cfg{288, 16, 16, 3}, Cfg{576, 16, 8, 2}, Cfg{576, 16, 12, 2}, Cfg{288, 16, 8, 4}, Cfg{288, 16, 8, 2}, Cfg{144, 16, 16, 4}}) {
        StreamParams p{src, total, c.S, c.rows, c.stages, sink};
        size_t sm = (size_t)c.S * c.rows * c.stages * c.warps;
        if (sm > g_smem_limit) continue;
        double gbs = timeit([&]() { stream_cpasync_kernel<<<g_nsm, c.warps * 32, sm>>>(p); });
        printf("cpasync S=%-5d rows=%-2d w=%-2d st=%d (ring %3zuK/SM): %7.1f\n", c.S, c.rows, c.warps, c.stages, sm / 1024, gbs);
    }
    for (int warps : {8, 16}) {
        StreamParams p{src, total, 0, 0, 0, sink};
        printf("ldg128  w=%-2d U=4: %7.1f   U=8: %7.1f   U=16: %7.1f\n", warps, timeit([&]() { stream_ldg_kernel<4><<<g_nsm, warps * 32>>>(p); }),
               timeit([&]() { stream_ldg_kernel<8><<<g_nsm, warps * 32>>>(p); }), timeit([&]() { stream_ldg_kernel<16><<<g_nsm, warps * 32>>>(p); }));
        printf("ldg128  w=%-2d 2 CTAs/SM U=8: %7.1f\n", warps, timeit([&]() { stream_ldg_kernel<8><<<2 * g_nsm, warps * 32>>>(p); }));
    }
    cudaFree(src); cudaFree(sink);
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "all";
    lab_init();
    if (mode == "check" || mode == "all") {
        Knobs k512{16, 3}, k1024{8, 2};
        for (int type : {T_Q4_K, T_Q5_K, T_Q6_K, T_Q8_0}) {
            check_onehot(type, 512, k512);
            check_onehot(type, 768, k1024);
        }
        check_onehot(T_Q8_0, 96, k512);
        for (const Knobs& kn : {k512, k1024, Knobs{12, 4}, Knobs{5, 2}}) {
            for (int type : {T_Q4_K, T_Q5_K, T_Q6_K, T_Q8_0}) {
                check_random(type, 4096, 4096, kn, ME_STORE, false, false);
                check_random(type, 2048, 1000, kn, ME_RESIDUAL, true, true);
            }
            check_random(T_Q4_K, 4096, 14336, kn, ME_SWIGLU, false, true);
            check_random(T_Q6_K, 14336, 4096, kn, ME_RESIDUAL, false, false);
            check_random(T_Q6_K, 5632, 2048, kn, ME_RESIDUAL, false, false);   // row bytes 4620: only 4-byte aligned rows
            check_random(T_Q8_0, 896, 1536, kn, ME_STORE, true, true);         // row bytes 952: 8-byte aligned rows, ragged last chunk
            check_random(T_Q5_K, 4864, 896, kn, ME_RESIDUAL, false, false);
        }
        check_random(T_Q4_K, 4096, 16, k512, ME_STORE, false, false);
        check_random(T_Q4_K, 256, 5, k512, ME_STORE, false, false);
        check_random(T_Q6_K, 4096, 128256, k512, ME_STORE, false, true);
        check_random(T_Q4_K, 4096, 4096, k512, ME_STORE, false, false, 200.0f);   // large activations
        check_random(T_Q4_K, 4096, 4096, k512, ME_STORE, false, false, 1e-4f);   // tiny activations
        check_qkv(k512);
        printf("CHECK SUMMARY: %s (%d failing groups)\n", g_fail ? "FAIL" : "ALL OK", g_fail);
    }
    if (mode == "rows") stream_rows_probe();
    if (mode == "stream" || mode == "all") stream_probes();
    if (mode == "prof") {   // one case, for ncu: gemv_lab prof <warps> <stages>
        Knobs kn{argc > 2 ? atoi(argv[2]) : 16, argc > 3 ? atoi(argv[3]) : 3};
        time_case({"gate/up swiglu 8B", T_Q4_K, 4096, 14336, ME_SWIGLU}, {kn}, false);
    }
    if (mode == "timeline") {
        Knobs kn{argc > 2 ? atoi(argv[2]) : 16, argc > 3 ? atoi(argv[3]) : 3};
        timeline_case({"O proj 8B", T_Q4_K, 4096, 4096, ME_RESIDUAL}, kn);
        timeline_case({"gate/up swiglu 8B", T_Q4_K, 4096, 14336, ME_SWIGLU}, kn);
        timeline_case({"down 8B Q4_K", T_Q4_K, 14336, 4096, ME_RESIDUAL}, kn);
    }
    if (mode == "chain") {   // gemv_lab chain <warps> <stages>: the PDL layer chain alone (for ncu launch lists)
        Knobs kn{argc > 2 ? atoi(argv[2]) : 16, argc > 3 ? atoi(argv[3]) : 3};
        time_layer_chain(kn, T_Q4_K, false);
        time_layer_chain(kn, T_Q4_K, true);
    }
    if (mode == "time" || mode == "all") {
        std::vector<Knobs> sweep = {{16, 3}, {16, 2}, {16, 4}, {12, 4}, {12, 3}, {8, 4}};
        std::vector<Knobs> one = {{16, 3}, {16, 2}, {12, 4}};
        time_case({"gate/up swiglu 8B", T_Q4_K, 4096, 14336, ME_SWIGLU}, sweep, true);
        time_case({"down 8B Q4_K", T_Q4_K, 14336, 4096, ME_RESIDUAL}, sweep, true);
        time_case({"down 8B Q6_K", T_Q6_K, 14336, 4096, ME_RESIDUAL}, sweep, true);
        time_case({"O proj 8B", T_Q4_K, 4096, 4096, ME_RESIDUAL}, sweep, true);
        time_case({"head 8B Q6_K", T_Q6_K, 4096, 128256, ME_STORE}, one, true);
        time_case({"Q5_K 4096x14336", T_Q5_K, 4096, 14336, ME_STORE}, one, true);
        time_case({"Q8_0 2048x5632", T_Q8_0, 2048, 5632, ME_STORE}, one, true);
        for (const Knobs& kn : {Knobs{16, 3}, Knobs{16, 2}, Knobs{12, 4}}) {
            time_layer_chain(kn, T_Q4_K, false);
            time_layer_chain(kn, T_Q4_K, true);
            time_layer_chain(kn, T_Q6_K, true);
        }
    }
    return g_fail ? 1 : 0;
}
