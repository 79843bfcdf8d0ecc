// engine.cu — context, weight residency, the per-token launch sequence and the C ABI
// (include/llama_b200.h) of the cuda-b200 backend.
//
// Mirrors GpuOnlyInference (src/backend/cuda/gpu_only.rs:24-88, 425-2011): weights are
// uploaded once and stay resident in their GGUF block layout; a token is one replay of a
// CUDA graph (embedding row -> L x [norm+QKV GEMV, RoPE+KV write, GQA attention,
// O GEMV+residual, norm+gate/up GEMV+SwiGLU, down GEMV+residual] -> norm+vocab GEMV).
// Only two host<->device crossings per token remain: 4 bytes of token id in, `vocab`
// f32 logits out (none at all in b200_decode_greedy).  Kernel-to-kernel edges are
// programmatic dependent launches so the next kernel's weight prefetch overlaps the
// tail of the previous one.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/llama_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "gemm_umma.cuh"
#include "gemm_umma2.cuh"
#include "attn_umma.cuh"
#include "kv_int8.cuh"
#include "gguf_load.cuh"
#include "gemv_mma.cuh"
#include "mega.cuh"
#include "misc.cuh"
#include "prefill.cuh"
#include "quant.cuh"
#include "stream.cuh"
#include "stream2.cuh"

using namespace b200;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(B200_ERR_OPERATION_FAILED, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define CU_ALLOC(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(B200_ERR_ALLOCATION_FAILED, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* b200_backend_name(void) { return "cuda-b200"; }
extern "C" const char* b200_last_error(void) { return g_err.c_str(); }
extern "C" int b200_device_count(int* out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (out) *out = 0;
        return fail(B200_ERR_NOT_AVAILABLE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    if (out) *out = n;
    return B200_OK;
}
extern "C" int b200_type_block_elems(uint32_t t) { return type_block_elems((int)t); }
extern "C" int b200_type_block_bytes(uint32_t t) { return type_block_bytes((int)t); }

static bool type_supported(int t) { return type_block_elems(t) != 0; }

// ------------------------------------------------------------------ context
struct DevTensor {
    uint8_t* d = nullptr;
    int type = 0;
    int n_dims = 0;
    uint64_t ne[4] = {1, 1, 1, 1};
    size_t nbytes = 0;
    long long row_bytes = 0;  // bytes of one output row (ne[0] elements)
    bool arena = false;       // d points into the context's load arena (b200_ctx_load_gguf): not freed on its own
    bool present() const { return d != nullptr; }
    const float* f32() const { return reinterpret_cast<const float*>(d); }
};

struct Layer {
    DevTensor attn_norm, ffn_norm, wq, wk, wv, wo, bq, bk, bv, gate, up, down;
    DevTensor router, gate_exps, up_exps, down_exps;
};

enum Mode { MODE_LOGITS = 0, MODE_PREFILL = 1, MODE_GREEDY = 2, MODE_COUNT = 3 };

struct Slot {
    SeqState* d_state = nullptr;
    int* d_generated = nullptr;
    float* kv = nullptr;  // [L][2][n_kv][max_seq][hd]
    signed char* kv8 = nullptr;   // int8 KV format (kv_int8.cuh): [L][2][n_kv][max_seq][hd] bytes ...
    float* kv_scale = nullptr;    // ... and [L][2][n_kv][max_seq] scales; kv stays null
    uint64_t host_pos = 0;
    bool spec = false;            // greedy continuation (b200_ctx_set_speculation): the NEXT token is already in flight ...
    int spec_buf = 0;             // ... its logits land in h_spec_logits[spec_buf], the token it consumed in h_spec_pick[spec_buf]
    cudaGraphExec_t graph[MODE_COUNT] = {nullptr, nullptr, nullptr};
    uint64_t graph_launches[MODE_COUNT] = {0, 0, 0};
    MegaPhase* d_phases = nullptr;  // per-token megakernel program of this slot (mega.cuh)
    MegaPhase* d_phases2 = nullptr; // the same program for stream2.cuh: EMBED phase first
    std::vector<uint32_t> pending;  // tokens queued by b200_prefill_token (B200_PREFILL_QUEUE=1), not yet run
};

constexpr int kMaxGenerated = 1 << 16;

struct b200_ctx {
    b200_model_desc d{};
    b200_parallel_desc par{1, 0, 0};
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    std::map<std::string, DevTensor> tensors;
    std::vector<Layer> layers;
    DevTensor token_embd, output_norm, output;
    bool finalized = false;
    // scratch
    float *xa = nullptr, *xb = nullptr, *qkv = nullptr, *attn = nullptr, *hbuf = nullptr, *logits = nullptr;
    float *attn_part = nullptr, *rope_freq = nullptr, *taps = nullptr, *moe_wt = nullptr;
    int* moe_sel = nullptr;
    // expert parallel (MoE model, world_size > 1): experts [rank * ep_local, (rank + 1) * ep_local) are resident here; everything else is
    // replicated; the weighted expert outputs are combined through (value, epoch) packets in the peers' exchange regions (misc.cuh)
    bool ep = false;
    int ep_local = 0;
    float* ep_y = nullptr;             // [8][hidden] weighted outputs of the local selected experts
    unsigned int* ep_epoch = nullptr;  // device counter, bumped by the routing kernel once per layer
    unsigned int* tickets = nullptr;
    int n_splits = 1;
    // tensor-pipe GEMV (gemv_mma.cuh): stream-K scratch, watchdog flag, launch knobs
    float* mma_part = nullptr;
    unsigned int* mma_tickets = nullptr;
    int* mma_err = nullptr;
    int mma_tickets_n = 0;
    bool use_mma = true;
    int mma_warps = 16, mma_stages = 3;
    size_t smem_optin = 227 * 1024;
    uint64_t mma_launches = 0, v1_launches = 0;
    // per-token megakernel (mega.cuh)
    // GGUF -> HBM load (gguf_load.cuh): two pinned staging buffers, worker threads fill one while the copy engine drains the other
    struct LoadStage {
        bool on = false;
        uint8_t* pin[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaStream_t st = nullptr;
        size_t chunk = 0;
        int next = 0, threads = 4;
        uint64_t bytes = 0;
    } stage;
    // one allocation for every tensor of a GGUF load (291 cudaMalloc + memset pairs cost more than copying Llama-3-8B's 4.9 GB):
    // zeroed once, tensors carved out at 256-byte boundaries with the 256 zero bytes behind each that the kernels' tail reads expect
    uint8_t* arena = nullptr;
    size_t arena_size = 0, arena_off = 0;
    // Greedy continuation behind b200_forward (opt-in): after the logits of token t the device picks argmax (last maximum wins) and
    // starts token t + 1 at once; if the caller's next token is that pick -- every greedy caller -- its logits are already on the way
    // and the host's turnaround (D2H, argmax, call overhead) overlaps the next token instead of idling the GPU.
    bool speculate = false;
    float* h_spec_logits[2] = {nullptr, nullptr};
    int* h_spec_pick = nullptr;   // pinned [2]
    cudaEvent_t spec_ev[2] = {nullptr, nullptr};
    uint64_t spec_hits = 0, spec_misses = 0;
    bool kv_int8_gemm = false;    // B200_KV_INT8_GEMM=1: int8 contexts take the tensor-core prompt path (see prefill_gemm_ok)
    int kv_format = 0;            // 0 = f32 (model/mod.rs:83-108), 1 = int8 (model/kv_quantized.rs Int8): B200_KV_FORMAT / b200_ctx_set_kv_format
    float* attn_q8_part = nullptr;
    int attn_q8_splits = 0;
    bool use_mega = true, mega_ok = false;
    int mega_phases = 0;
    size_t mega_smem = 0;
    unsigned int* mega_bar = nullptr;
    float* mega_cand_val = nullptr;
    int* mega_cand_idx = nullptr;
    float* mega_attn_part = nullptr;
    uint8_t* mega_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // staged (int8 planes) forms of xa, xb, attn, hbuf
    int mega_splits = 1;
    uint64_t mega_launches = 0;
    unsigned long long* mega_dbg = nullptr;
    // streamed megakernel (stream.cuh): TMA tensor maps of the weight matrices, ring geometry
    bool use_stream = true, stream_ok = false;
    // second streamed megakernel (stream2.cuh)
    bool use_stream2 = true, stream2_ok = false;
    int s2_phases = 0, s2_xr_off = 0, s2_tpart_off = 0, s2_desc_off = 0, s2_ring_off = 0, s2_slots = 0;
    size_t s2_smem = 0;
    uint2* s2_ll = nullptr;
    unsigned int s2_epoch = 0;
    float* s2_cand_val = nullptr;
    int* s2_cand_idx = nullptr;
    void* d_tmaps = nullptr;
    size_t stream_smem = 0;
    int stream_ring_off = 0, stream_slots = 0;
    // tensor parallel (one context per rank/GPU; c->d holds the LOCAL head / ffn counts, dg the global ones)
    b200_model_desc dg{};
    int vocab_l = 0;                     // rows of the vocab head owned by this rank
    uint8_t* tp_region = nullptr;        // [2][P][H] f32 partial sums | flags[16] u32 | cand[P][2] f32   (IPC-exported)
    uint8_t* tp_peer[kMmaMaxPeers] = {}; // the same region of every rank as mapped in this process (own rank: tp_region)
    bool tp_peer_set[kMmaMaxPeers] = {};
    unsigned int tp_epoch = 0;
    size_t out_scratch_elems = 0;
    // pinned host staging
    float* h_logits = nullptr;
    // GEMM prefill (gemm_umma.cuh + prefill.cuh): activation rows of one chunk of prompt tokens
    bool use_prefill_gemm = true;
    int prefill_gemm_min = 32;
    bool queue_bypass = false;    // (set while a queue is being flushed: its short tail runs token by token)
    bool prefill_queue = false;   // b200_prefill_token queues; the queue runs as ONE GEMM prefill at the next call that needs the state
    float* pf_buf = nullptr;
    int* pf_tok = nullptr;
    uint8_t* pf_rows = nullptr;      // batched decode: per-row position | KV base pointer | SeqState pointer
    float* pf_logits = nullptr;      // batched decode: [rows][vocab]
    float* pf_split = nullptr;       // split-K partial tiles of the small-T GEMMs
    __half* pf_k16 = nullptr;        // fp16 copies of one layer's K rows / transposed V rows for the tensor-core prefill attention
    __half* pf_vt16 = nullptr;
    int pf_kv16_P = 0;
    CUtensorMap pf_kmap, pf_vmap;
    bool pf_attn_tc = true;          // B200_PREFILL_ATTN_TC=0: the CUDA-core prefill attention
    int* pf_tile_cnt = nullptr;      // split-K tile counters of the persistent dequant-GEMM (zeroed once; the kernel re-zeroes them)
    void* pf_tmaps = nullptr;        // TMA tensor maps (128-row boxes) of the GEMM weights
    std::map<const void*, int> pf_tmap_of;
    bool pf_tmaps_built = false;
    size_t pf_split_floats = 0;
    int pf_logits_rows = 0;
    // batched decode: the launches of a step are the same for a given row count (positions, cache bases and slot states are read from
    // device arrays) -> captured once per row count into a CUDA graph and replayed (B200_BATCH_GRAPH=0: eager launches)
    struct BatchGraph { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; bool warm = false; };
    std::vector<BatchGraph> batch_graphs;
    bool batch_graph = true;
    bool batch_pdl = true;            // ... with programmatic dependent launches between its kernels (B200_BATCH_PDL=0: plain edges)
    float* h_batch_logits = nullptr; // pinned staging of b200_decode_batch's n x vocab logits
    size_t h_batch_logits_bytes = 0;
    int* pf_argmax = nullptr;        // batched decode with the pick on the device: [rows]
    int pf_argmax_rows = 0;
    int batch_gemm_min = 8;    // measured crossover on Llama-3-8B: a GEMM pass costs ~12 ms up to 32 rows, a sequence alone 2 ms
    uint64_t prefill_gemm_tokens = 0;
    int* h_err = nullptr;   // pinned copy of the first watchdog word, fetched with every synchronising call
    int* h_token = nullptr;
    std::vector<Slot> slots;
    // options
    bool use_graph = true, use_pdl = true, use_taps = false;
    // statistics
    uint64_t launches = 0;
    uint64_t weight_bytes_per_token = 0;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    cudaGraphExec_t gemv_graph = nullptr;  // b200_bench_gemv_pass
    uint64_t gemv_graph_launches = 0;
};

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_k(b200_ctx* c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = c->use_pdl ? at : nullptr;
    cfg.numAttrs = c->use_pdl ? 1 : 0;
    c->launches++;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

extern "C" int b200_ctx_create(const b200_model_desc* desc, const b200_parallel_desc* par, b200_ctx** out) {
    if (!desc || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_create: null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(B200_ERR_NOT_AVAILABLE, "cuda-b200: no CUDA device (this backend has no CPU fallback)");
    }
    b200_ctx* c = new b200_ctx();
    c->d = *desc;
    if (par) c->par = *par;
    b200_model_desc& d = c->d;
    if (d.head_dim <= 0) d.head_dim = d.hidden / std::max(d.n_heads, 1);
    if (d.rope_scale == 0.0f) d.rope_scale = 1.0f;
    if (d.max_batch <= 0) d.max_batch = 1;
    if (d.n_experts > 0 && d.expert_ffn <= 0) d.expert_ffn = d.ffn;
    auto bad = [&](const char* m) {
        delete c;
        return fail(B200_ERR_INVALID_ARGUMENT, std::string("b200_ctx_create: ") + m);
    };
    if (d.hidden <= 0 || d.n_layers <= 0 || d.n_heads <= 0 || d.n_kv_heads <= 0 || d.vocab <= 0 || d.max_seq_len <= 0)
        return bad("non-positive model dimension");
    if (d.n_heads % d.n_kv_heads) return bad("n_heads must be a multiple of n_kv_heads");
    if (d.head_dim != 64 && d.head_dim != 128) return bad("head_dim must be 64 or 128");
    if (d.n_heads / d.n_kv_heads > 8) return bad("more than 8 query heads per kv head");
    if (d.hidden % 32) return bad("hidden must be a multiple of 32");
    if (d.n_experts > 64 || d.n_experts_used > 8) return bad("at most 64 experts, top-8");
    if (c->par.world_size < 1 || c->par.rank < 0 || c->par.rank >= c->par.world_size) return bad("bad parallel desc");
    c->dg = d;
    if (c->par.world_size != 1) {
        const int P = c->par.world_size;
        if (P != 2 && P != 4 && P != 8) return bad("tensor-parallel world size must be 2, 4 or 8");
        if (d.n_experts > 0) {
            // MoE: EXPERT parallel -- attention, norms, router and the vocab head replicated, expert e on rank e / (E / P); at batch 1 the
            // dispatch is free (every rank routes), the combine is the only exchange (replaces moe.rs:352-361 on one host)
            if (d.n_experts % P) return bad("n_experts must be divisible by the world size");
            c->ep = true;
            c->ep_local = d.n_experts / P;
        } else {
        // ShardingPlan::from_config's divisibility rules (src/backend/tensor_parallel.rs:69-106)
        if (d.n_heads % P || d.n_kv_heads % P || d.ffn % P || d.vocab % P) return bad("heads, kv heads, ffn and vocab must be divisible by the world size");
        if ((d.vocab / P) % 16) return bad("vocab / world_size must be a multiple of 16");
        d.n_heads /= P;
        d.n_kv_heads /= P;
        d.ffn /= P;
        }
    }
    c->vocab_l = c->ep ? d.vocab : d.vocab / c->par.world_size;
    if (c->par.device >= n_dev) return bad("device ordinal out of range");
    CU(cudaSetDevice(c->par.device));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, c->par.device));
    c->n_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->use_graph = env_int("B200_GRAPH", 1) != 0;
    c->use_pdl = env_int("B200_PDL", 1) != 0;
    c->use_taps = env_int("B200_TAPS", 0) != 0;
    c->use_mma = env_int("B200_GEMV_MMA", 1) != 0;
    c->use_mega = env_int("B200_MEGA", 1) != 0;
    c->use_stream = env_int("B200_STREAM", 1) != 0;
    c->use_stream2 = env_int("B200_STREAM2", 1) != 0;
    c->use_prefill_gemm = env_int("B200_PREFILL_GEMM", 1) != 0;
    c->pf_attn_tc = env_int("B200_PREFILL_ATTN_TC", 1) != 0;
    {
        const char* kf = getenv("B200_KV_FORMAT");
        if (kf && (!strcmp(kf, "int8") || !strcmp(kf, "1"))) c->kv_format = 1;
        c->kv_int8_gemm = env_int("B200_KV_INT8_GEMM", 0) != 0;
    }
    c->prefill_gemm_min = std::max(1, env_int("B200_PREFILL_GEMM_MIN", 32));
    c->prefill_queue = env_int("B200_PREFILL_QUEUE", 0) != 0;
    c->batch_gemm_min = std::max(2, env_int("B200_BATCH_GEMM_MIN", 8));
    c->batch_graph = env_int("B200_BATCH_GRAPH", 1) != 0;
    c->batch_pdl = env_int("B200_BATCH_PDL", 1) != 0;
    c->mma_warps = std::max(4, std::min(kMmaMaxWarps, env_int("B200_MMA_WARPS", 16)));
    c->mma_stages = std::max(2, std::min(kMmaMaxStages, env_int("B200_MMA_STAGES", 3)));
    c->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
    c->layers.resize(d.n_layers);
    *out = c;
    return B200_OK;
}

static DevTensor* slot_for_name(b200_ctx* c, const std::string& name) {
    if (name == "token_embd.weight") return &c->token_embd;
    if (name == "output_norm.weight") return &c->output_norm;
    if (name == "output.weight") return &c->output;
    int li = -1;
    char rest[96];
    if (sscanf(name.c_str(), "blk.%d.%95s", &li, rest) == 2 && li >= 0 && li < (int)c->layers.size()) {
        Layer& L = c->layers[li];
        std::string r(rest);
        if (r == "attn_norm.weight") return &L.attn_norm;
        if (r == "ffn_norm.weight") return &L.ffn_norm;
        if (r == "attn_q.weight") return &L.wq;
        if (r == "attn_k.weight") return &L.wk;
        if (r == "attn_v.weight") return &L.wv;
        if (r == "attn_output.weight") return &L.wo;
        if (r == "attn_q.bias") return &L.bq;
        if (r == "attn_k.bias") return &L.bk;
        if (r == "attn_v.bias") return &L.bv;
        if (r == "ffn_gate.weight") return &L.gate;
        if (r == "ffn_up.weight") return &L.up;
        if (r == "ffn_down.weight") return &L.down;
        if (r == "ffn_gate_inp.weight") return &L.router;
        if (r == "ffn_gate_exps.weight") return &L.gate_exps;
        if (r == "ffn_up_exps.weight") return &L.up_exps;
        if (r == "ffn_down_exps.weight") return &L.down_exps;
    }
    return nullptr;
}

// 0 = replicated, 1 = column-parallel (split output rows), 2 = row-parallel (split K)
static int tp_shard_kind(const char* name) {
    const std::string n(name);
    auto ends = [&](const char* suf) { const std::string s2(suf); return n.size() >= s2.size() && n.compare(n.size() - s2.size(), s2.size(), s2) == 0; };
    if (n == "output.weight") return 1;
    if (ends("attn_q.weight") || ends("attn_k.weight") || ends("attn_v.weight") || ends("attn_q.bias") || ends("attn_k.bias") ||
        ends("attn_v.bias") || ends("ffn_gate.weight") || ends("ffn_up.weight"))
        return 1;
    if (ends("attn_output.weight") || ends("ffn_down.weight")) return 2;
    return 0;
}

// Device memory of one tensor: `bytes` + 256 zero bytes behind it.  Out of the load arena while one is open (already zeroed),
// else its own allocation; `zero_body` also clears the tensor's bytes (row-parallel shards: the pitch gaps must read as zero).
static cudaError_t tensor_alloc(b200_ctx* c, DevTensor* t, size_t bytes, bool zero_body) {
    const size_t need = (bytes + 256 + 255) & ~(size_t)255;
    if (c->arena && c->arena_off + need <= c->arena_size) {
        t->d = c->arena + c->arena_off;
        t->arena = true;
        c->arena_off += need;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc((void**)&t->d, bytes + 256);
    if (e != cudaSuccess) return e;
    t->arena = false;
    return zero_body ? cudaMemset(t->d, 0, bytes + 256) : cudaMemset(t->d + bytes, 0, 256);
}

// Host -> device copies of b200_ctx_upload_tensor.  Plain cudaMemcpy for caller-owned buffers; while a GGUF load is running
// (c->stage.on) the bytes come out of a file mapping and go through the pinned double buffer instead: fill buffer b with host
// threads, cudaMemcpyAsync it on the load stream, and meanwhile fill buffer b ^ 1.  The load stream is a BLOCKING stream, so the
// legacy-stream memsets of the allocation are ordered with the copies.
static cudaError_t h2d(b200_ctx* c, void* dst, const void* src, size_t n) {
    b200_ctx::LoadStage& s = c->stage;
    if (!s.on) {
        s.bytes += n;   // (b200_load_stats.device_bytes of an unstaged load)
        return cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
    }
    cudaError_t e;
    for (size_t off = 0; off < n; off += s.chunk) {
        const size_t len = std::min(s.chunk, n - off);
        const int b = s.next;
        s.next ^= 1;
        if ((e = cudaEventSynchronize(s.ev[b])) != cudaSuccess) return e;
        gguf_parallel_copy(s.pin[b], (const uint8_t*)src + off, len, s.threads);
        if ((e = cudaMemcpyAsync((uint8_t*)dst + off, s.pin[b], len, cudaMemcpyHostToDevice, s.st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(s.ev[b], s.st)) != cudaSuccess) return e;
        s.bytes += len;
    }
    return cudaSuccess;
}
// rows of `width` bytes, `spitch` apart on the host, `dpitch` apart on the device (row-parallel shards: a K range of every row)
static cudaError_t h2d_2d(b200_ctx* c, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows) {
    b200_ctx::LoadStage& s = c->stage;
    if (!s.on) {
        s.bytes += width * rows;
        return cudaMemcpy2D(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice);
    }
    cudaError_t e;
    const size_t rpc = std::max<size_t>(1, s.chunk / width);
    if (width > s.chunk) return cudaErrorInvalidValue;
    for (size_t r0 = 0; r0 < rows; r0 += rpc) {
        const size_t nr = std::min(rpc, rows - r0);
        const int b = s.next;
        s.next ^= 1;
        if ((e = cudaEventSynchronize(s.ev[b])) != cudaSuccess) return e;
        uint8_t* pin = s.pin[b];
        const uint8_t* from = (const uint8_t*)src + r0 * spitch;
        const int nt = (int)std::min<size_t>((size_t)std::max(1, s.threads), nr);
        std::vector<std::thread> th;
        auto gather = [=](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) memcpy(pin + i * width, from + i * spitch, width); };
        const size_t per = (nr + nt - 1) / nt;
        for (int t = 1; t < nt; t++)
            if (per * t < nr) th.emplace_back(gather, per * t, std::min(nr, per * (t + 1)));
        gather(0, std::min(nr, per));
        for (std::thread& t : th) t.join();
        if ((e = cudaMemcpy2DAsync((uint8_t*)dst + r0 * dpitch, dpitch, pin, width, width, nr, cudaMemcpyHostToDevice, s.st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(s.ev[b], s.st)) != cudaSuccess) return e;
        s.bytes += nr * width;
    }
    return cudaSuccess;
}

extern "C" int b200_ctx_upload_tensor(b200_ctx* c, const char* gguf_name, uint32_t ggml_type, const uint64_t* ne,
                                      int n_dims, const void* host, size_t nbytes) {
    if (!c || !gguf_name || !ne || !host) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_upload_tensor: null argument");
    if (c->finalized) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_upload_tensor: context already finalized");
    if (n_dims < 1 || n_dims > 3) return fail(B200_ERR_SHAPE_MISMATCH, std::string(gguf_name) + ": 1..3 dims expected");
    if (!type_supported((int)ggml_type))
        return fail(B200_ERR_UNSUPPORTED_DTYPE, std::string(gguf_name) + ": unsupported ggml type " + std::to_string(ggml_type));
    DevTensor* t = slot_for_name(c, gguf_name);
    if (!t) return fail(B200_ERR_INVALID_ARGUMENT, std::string("unknown tensor name: ") + gguf_name);
    const int be = type_block_elems((int)ggml_type), bb = type_block_bytes((int)ggml_type);
    uint64_t numel = 1;
    for (int i = 0; i < n_dims; i++) numel *= ne[i];
    if (ne[0] % be) return fail(B200_ERR_SHAPE_MISMATCH, std::string(gguf_name) + ": ne[0] not a multiple of the block size");
    if (numel / be * bb != nbytes) return fail(B200_ERR_SHAPE_MISMATCH, std::string(gguf_name) + ": byte size does not match shape");
    CU(cudaSetDevice(c->par.device));
    if (t->d && !t->arena) cudaFree(t->d);
    *t = DevTensor();
    // Tensor-parallel shard of this rank (Megatron split; ShardingPlan, src/backend/tensor_parallel.rs:69-106):
    //   column-parallel = a contiguous range of output rows (attn_q/k/v + biases by head, ffn_gate/up, output),
    //   row-parallel    = a contiguous range of K blocks of EVERY row (attn_output, ffn_down),
    //   everything else replicated.  GGUF rows are contiguous blocks, so both are plain (2-D) byte ranges.
    const int P = c->par.world_size, R = c->par.rank;
    uint64_t lne[4] = {1, 1, 1, 1};
    for (int i = 0; i < n_dims; i++) lne[i] = ne[i];
    const uint64_t row_bytes_full = ne[0] / be * bb;
    const bool is_exps = std::string(gguf_name).find("_exps.weight") != std::string::npos;
    const int kind = (P > 1 && c->ep) ? (is_exps ? 3 : 0) : (P > 1) ? tp_shard_kind(gguf_name) : 0;
    long long tp_pitch = 0;
    if (kind == 1) {
        const int dim = n_dims == 1 ? 0 : 1;
        if (ne[dim] % P) return fail(B200_ERR_SHAPE_MISMATCH, std::string(gguf_name) + ": rows not divisible by the world size");
        lne[dim] = ne[dim] / P;
        const size_t lbytes = nbytes / P;
        CU_ALLOC(tensor_alloc(c, t, lbytes, false));
        CU(h2d(c, t->d, (const uint8_t*)host + (size_t)R * lbytes, lbytes));
        t->nbytes = lbytes;
    } else if (kind == 3) {   // expert parallel: the experts are the outermost dimension -> a contiguous byte range per rank
        if (n_dims != 3 || ne[2] % P) return fail(B200_ERR_SHAPE_MISMATCH, std::string(gguf_name) + ": experts not divisible by the world size");
        lne[2] = ne[2] / P;
        const size_t lbytes = nbytes / P;
        CU_ALLOC(tensor_alloc(c, t, lbytes, false));
        CU(h2d(c, t->d, (const uint8_t*)host + (size_t)R * lbytes, lbytes));
        t->nbytes = lbytes;
    } else if (kind == 2) {
        const uint64_t nb = ne[0] / be;
        if (nb % P) return fail(B200_ERR_SHAPE_MISMATCH, std::string(gguf_name) + ": K blocks not divisible by the world size");
        lne[0] = ne[0] / P;
        // rows keep their bytes, but start 16 bytes apart at least: a row pitch that is a 16-byte multiple is what the TMA
        // tensor map of the streamed kernel needs (Llama-3-8B TP=2: 28 Q6_K blocks = 5880 bytes -> pitch 5888)
        const size_t slice = (size_t)(nb / P) * bb, rows = (size_t)(numel / ne[0]);
        const size_t pitch = (slice + 15) & ~(size_t)15;
        CU_ALLOC(tensor_alloc(c, t, pitch * rows, true));
        CU(h2d_2d(c, t->d, pitch, (const uint8_t*)host + (size_t)R * slice, row_bytes_full, slice, rows));
        t->nbytes = pitch * rows;
        tp_pitch = (long long)pitch;
    } else {
        CU_ALLOC(tensor_alloc(c, t, nbytes, false));
        CU(h2d(c, t->d, host, nbytes));
        t->nbytes = nbytes;
    }
    t->type = (int)ggml_type;
    t->n_dims = n_dims;
    for (int i = 0; i < n_dims; i++) t->ne[i] = lne[i];
    t->row_bytes = tp_pitch ? tp_pitch : (long long)(lne[0] / be * bb);
    c->tensors[gguf_name] = *t;
    return B200_OK;
}

static int mega_build(b200_ctx* c);
static int check_weight(const DevTensor& t, const char* what, uint64_t k, uint64_t n, int layer) {
    std::string nm = std::string(what) + (layer >= 0 ? " (layer " + std::to_string(layer) + ")" : "");
    if (!t.present()) return fail(B200_ERR_INVALID_ARGUMENT, "missing tensor " + nm);
    if (t.ne[0] != k || t.ne[1] != n)
        return fail(B200_ERR_SHAPE_MISMATCH, nm + ": expected [" + std::to_string(k) + "," + std::to_string(n) + "], got [" +
                                                 std::to_string(t.ne[0]) + "," + std::to_string(t.ne[1]) + "]");
    if (k % 32) return fail(B200_ERR_SHAPE_MISMATCH, nm + ": in_features must be a multiple of 32");
    return B200_OK;
}
static int check_f32_vec(const DevTensor& t, const char* what, uint64_t n, int layer, bool required) {
    std::string nm = std::string(what) + (layer >= 0 ? " (layer " + std::to_string(layer) + ")" : "");
    if (!t.present()) return required ? fail(B200_ERR_INVALID_ARGUMENT, "missing tensor " + nm) : B200_OK;
    if (t.type != T_F32) return fail(B200_ERR_DTYPE_MISMATCH, nm + ": must be F32");
    if (t.ne[0] != n) return fail(B200_ERR_SHAPE_MISMATCH, nm + ": wrong length");
    return B200_OK;
}

extern "C" int b200_ctx_finalize(b200_ctx* c) {
    if (!c) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_finalize: null ctx");
    if (c->finalized) return B200_OK;
    const b200_model_desc& d = c->d;
    CU(cudaSetDevice(c->par.device));
    const uint64_t H = d.hidden, hd = d.head_dim, nh = d.n_heads, nkv = d.n_kv_heads, I = d.ffn, V = d.vocab;
    int rc;
    if ((rc = check_weight(c->token_embd, "token_embd.weight", H, V, -1))) return rc;
    if ((rc = check_f32_vec(c->output_norm, "output_norm.weight", H, -1, true))) return rc;
    if (c->output.present()) {
        if ((rc = check_weight(c->output, "output.weight", H, (uint64_t)c->vocab_l, -1))) return rc;
    } else if (!d.tied_output) {
        // the loader ties silently when output.weight is absent (loader.rs:349-355)
    }
    uint64_t wbytes = c->token_embd.row_bytes;  // one embedding row per token
    wbytes += (c->output.present() ? c->output.nbytes : c->token_embd.nbytes) + H * 4;
    for (int l = 0; l < d.n_layers; l++) {
        Layer& L = c->layers[l];
        if ((rc = check_f32_vec(L.attn_norm, "attn_norm.weight", H, l, true))) return rc;
        if ((rc = check_f32_vec(L.ffn_norm, "ffn_norm.weight", H, l, true))) return rc;
        if ((rc = check_weight(L.wq, "attn_q.weight", H, nh * hd, l))) return rc;
        if ((rc = check_weight(L.wk, "attn_k.weight", H, nkv * hd, l))) return rc;
        if ((rc = check_weight(L.wv, "attn_v.weight", H, nkv * hd, l))) return rc;
        if ((rc = check_weight(L.wo, "attn_output.weight", nh * hd, H, l))) return rc;
        if ((rc = check_f32_vec(L.bq, "attn_q.bias", nh * hd, l, false))) return rc;
        if ((rc = check_f32_vec(L.bk, "attn_k.bias", nkv * hd, l, false))) return rc;
        if ((rc = check_f32_vec(L.bv, "attn_v.bias", nkv * hd, l, false))) return rc;
        wbytes += L.wq.nbytes + L.wk.nbytes + L.wv.nbytes + L.wo.nbytes + 2 * H * 4;
        if (d.n_experts > 0) {
            const uint64_t EI = d.expert_ffn, E = d.n_experts, EL = c->ep ? (uint64_t)c->ep_local : E;   // (EL: experts resident on this GPU)
            if (!L.router.present() || L.router.type != T_F32 || L.router.ne[0] != H || L.router.ne[1] != E)
                return fail(B200_ERR_DTYPE_MISMATCH, "ffn_gate_inp.weight must be F32 [hidden, n_experts] (moe.rs:133)");
            for (const DevTensor* t : {&L.gate_exps, &L.up_exps}) {
                if ((rc = check_weight(*t, "ffn_{gate,up}_exps.weight", H, EI, l))) return rc;
                if (t->ne[2] != EL) return fail(B200_ERR_SHAPE_MISMATCH, "expert tensor: ne[2] != n_experts");
            }
            if ((rc = check_weight(L.down_exps, "ffn_down_exps.weight", EI, H, l))) return rc;
            if (L.down_exps.ne[2] != EL) return fail(B200_ERR_SHAPE_MISMATCH, "expert tensor: ne[2] != n_experts");
            if (L.gate_exps.type != L.up_exps.type)
                return fail(B200_ERR_UNSUPPORTED, "gate/up expert tensors must share one quant type");
            wbytes += L.router.nbytes +
                      (L.gate_exps.nbytes + L.up_exps.nbytes + L.down_exps.nbytes) / EL * d.n_experts_used;   // (expert parallel: summed over the GPUs)
        } else {
            if ((rc = check_weight(L.gate, "ffn_gate.weight", H, I, l))) return rc;
            if ((rc = check_weight(L.up, "ffn_up.weight", H, I, l))) return rc;
            if ((rc = check_weight(L.down, "ffn_down.weight", I, H, l))) return rc;
            if (L.gate.type != L.up.type) return fail(B200_ERR_UNSUPPORTED, "ffn_gate/ffn_up must share one quant type");
            wbytes += L.gate.nbytes + L.up.nbytes + L.down.nbytes;
        }
    }
    c->weight_bytes_per_token = wbytes;

    const uint64_t ffn_w = d.n_experts > 0 ? (uint64_t)d.expert_ffn : I;
    CU_ALLOC(cudaMalloc((void**)&c->xa, H * 4));
    CU_ALLOC(cudaMalloc((void**)&c->xb, H * 4));
    CU_ALLOC(cudaMalloc((void**)&c->qkv, (nh + 2 * nkv) * hd * 4));
    CU_ALLOC(cudaMalloc((void**)&c->attn, nh * hd * 4));
    CU_ALLOC(cudaMalloc((void**)&c->hbuf, ffn_w * 4));
    c->out_scratch_elems = std::max<uint64_t>(std::max<uint64_t>(V, ffn_w), std::max<uint64_t>(H, nh * hd));
    CU_ALLOC(cudaMalloc((void**)&c->logits, c->out_scratch_elems * 4));
    c->n_splits = (int)std::min<uint64_t>(64, std::max<uint64_t>(1, (2 * (uint64_t)c->n_sm + nkv - 1) / nkv));
    {   // the split merge stages n_splits * G * (hd + 2) floats in shared memory: it must fit the opt-in limit (G = 7..8 with hd = 128
        // and 64 splits would need 233-266 KB: Qwen2.5-3B / Qwen2-7B shapes)
        const uint64_t G = nh / nkv, lim = ((uint64_t)c->smem_optin - 2048) / (G * (hd + 2) * 4);
        c->n_splits = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->n_splits, lim));
    }
    CU_ALLOC(cudaMalloc((void**)&c->attn_part, nkv * c->n_splits * (nh / nkv) * (hd + 2) * 4));
    CU_ALLOC(cudaMalloc((void**)&c->tickets, nkv * sizeof(unsigned int)));
    CU(cudaMemset(c->tickets, 0, nkv * sizeof(unsigned int)));
    CU_ALLOC(cudaMalloc((void**)&c->mma_part, (size_t)c->n_sm * kMmaMaxWarps * 2 * 32 * sizeof(float)));
    c->mma_tickets_n = (int)((std::max<uint64_t>(std::max<uint64_t>(V, 2 * ffn_w), (nh + 2 * nkv) * hd) + 15) / 16 + 8);
    CU_ALLOC(cudaMalloc((void**)&c->mma_tickets, (size_t)c->mma_tickets_n * sizeof(unsigned int)));
    CU(cudaMemset(c->mma_tickets, 0, (size_t)c->mma_tickets_n * sizeof(unsigned int)));
    CU_ALLOC(cudaMalloc((void**)&c->mma_err, 8 * sizeof(int)));
    CU(cudaMemset(c->mma_err, 0, 8 * sizeof(int)));
    if (c->ep) {
        CU_ALLOC(cudaMalloc((void**)&c->ep_y, (size_t)8 * H * sizeof(float)));
        CU(cudaMemset(c->ep_y, 0, (size_t)8 * H * sizeof(float)));
        CU_ALLOC(cudaMalloc((void**)&c->ep_epoch, sizeof(unsigned int)));
        CU(cudaMemset(c->ep_epoch, 0, sizeof(unsigned int)));
    }
    CU_ALLOC(cudaMalloc((void**)&c->moe_sel, 8 * sizeof(int)));
    CU_ALLOC(cudaMalloc((void**)&c->moe_wt, 8 * sizeof(float)));
    CU(cudaMemset(c->moe_sel, 0, 8 * sizeof(int)));
    CU(cudaMemset(c->moe_wt, 0, 8 * sizeof(float)));
    CU_ALLOC(cudaMalloc((void**)&c->taps, (uint64_t)(d.n_layers + 1) * H * 4));
    CU(cudaMemset(c->taps, 0, (uint64_t)(d.n_layers + 1) * H * 4));
    // RoPE frequencies with the host's libm powf, exactly as the reference computes them
    // per element (cpu/ops.rs:1305): freq = 1 / base^(2i/hd).
    {
        std::vector<float> f(hd / 2);
        for (uint64_t i = 0; i < hd / 2; i++) f[i] = 1.0f / powf(d.rope_base, (float)(2 * i) / (float)hd);
        CU_ALLOC(cudaMalloc((void**)&c->rope_freq, f.size() * 4));
        CU(cudaMemcpy(c->rope_freq, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    }
    CU_ALLOC(cudaHostAlloc((void**)&c->h_logits, V * 4, cudaHostAllocDefault));
    CU_ALLOC(cudaHostAlloc((void**)&c->h_err, 8 * sizeof(int), cudaHostAllocDefault));
    memset(c->h_err, 0, 8 * sizeof(int));
    CU_ALLOC(cudaHostAlloc((void**)&c->h_token, sizeof(int), cudaHostAllocDefault));
    c->slots.resize(d.max_batch);
    const uint64_t kv_elems = (uint64_t)d.n_layers * 2 * nkv * d.max_seq_len * hd;
    for (Slot& s : c->slots) {
        CU_ALLOC(cudaMalloc((void**)&s.d_state, sizeof(SeqState)));
        CU(cudaMemset(s.d_state, 0, sizeof(SeqState)));
        CU_ALLOC(cudaMalloc((void**)&s.d_generated, kMaxGenerated * sizeof(int)));
        if (c->kv_format == 1) {   // int8 KV: bytes + one f32 scale per (kv head, position) row, no f32 cache at all
            CU_ALLOC(cudaMalloc((void**)&s.kv8, kv_elems));
            CU(cudaMemset(s.kv8, 0, kv_elems));
            CU_ALLOC(cudaMalloc((void**)&s.kv_scale, kv_elems / hd * 4));
            CU(cudaMemset(s.kv_scale, 0, kv_elems / hd * 4));
        } else {
            CU_ALLOC(cudaMalloc((void**)&s.kv, kv_elems * 4));
            CU(cudaMemset(s.kv, 0, kv_elems * 4));
        }
    }
    if (c->kv_format == 1) {
        if (c->par.world_size > 1) return fail(B200_ERR_UNSUPPORTED, "the int8 KV format is not available under tensor / expert parallelism");
        if (hd != 64 && hd != 128) return fail(B200_ERR_UNSUPPORTED, "the int8 KV format needs head_dim 64 or 128");
        if (nh % nkv || nh / nkv > 8) return fail(B200_ERR_UNSUPPORTED, "the int8 KV format needs at most 8 query heads per kv head");
        // per-op (graph) decode path only: the megakernels and the batched-decode pass read the f32 cache (the tensor-core PROMPT
        // path has int8 peers of its cache kernels: prefill_gemm)
        c->use_mega = false;
        c->attn_q8_splits = std::max(1, std::min(kAttnQ8MaxSplits, (2 * c->n_sm + (int)nkv - 1) / (int)nkv));   // two resident CTAs per SM
        CU_ALLOC(cudaMalloc((void**)&c->attn_q8_part, (size_t)nkv * c->attn_q8_splits * (nh / nkv) * (hd + 2) * sizeof(float)));
    }
    // kernels that may need more than 48 KB of dynamic shared memory
    CU(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CU(mma_set_smem_limit((int)c->smem_optin - 8192));
    {
        const int lim_attn = (int)c->smem_optin - 1024;
        CU(cudaFuncSetAttribute(attn_decode_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
        CU(cudaFuncSetAttribute(attn_decode_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
        CU(cudaFuncSetAttribute(attn_decode_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
        CU(cudaFuncSetAttribute(attn_decode_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
    }
    c->finalized = true;
    if (c->par.world_size > 1) {
        for (int r = 0; r < c->par.world_size; r++)
            if (!c->tp_peer_set[r]) {
                c->finalized = false;
                return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_finalize: tensor-parallel peers not connected (b200_ctx_tp_handle / b200_ctx_tp_set_peer)");
            }
    }
    if ((rc = mega_build(c))) return rc;
    if (c->par.world_size > 1 && !c->ep && !c->mega_ok) {
        c->finalized = false;
        return fail(B200_ERR_UNSUPPORTED, "tensor parallelism needs the megakernel path (dense model, K-quant / Q8_0 weights with 256-aligned shards)");
    }
    return B200_OK;
}

// ---- tensor-parallel plumbing: every rank exports one region (partial sums, flags, argmax candidates) as a CUDA IPC
// handle; the host side (torch.distributed all_gather in the Python mirror, MPI/gRPC in the Rust shim) passes the
// handles around; the kernels then read and write peer memory directly over NVLink.
static size_t tp_ll_off(const b200_ctx* c) {
    return ((size_t)2 * kMmaMaxPeers * c->d.hidden * sizeof(float) + 64 * sizeof(unsigned int) + kMmaMaxPeers * 2 * sizeof(float) + 255) & ~(size_t)255;
}
// ... | [2][P][H] (value, epoch) packets (stream2.cuh, ll_red: the all-reduce inside the row-parallel GEMV's phase)
static size_t tp_region_bytes(const b200_ctx* c) { return tp_ll_off(c) + (size_t)2 * kMmaMaxPeers * c->d.hidden * sizeof(uint2) + 256; }
static uint2* tp_ll(const b200_ctx* c, uint8_t* base, int buf, int rank_slot) {
    return reinterpret_cast<uint2*>(base + tp_ll_off(c)) + ((size_t)buf * kMmaMaxPeers + rank_slot) * c->d.hidden;
}
static float* tp_ar(const b200_ctx* c, uint8_t* base, int buf, int rank_slot) {
    return reinterpret_cast<float*>(base) + ((size_t)buf * kMmaMaxPeers + rank_slot) * c->d.hidden;
}
static unsigned int* tp_flags(const b200_ctx* c, uint8_t* base) {
    return reinterpret_cast<unsigned int*>(base + (size_t)2 * kMmaMaxPeers * c->d.hidden * sizeof(float));
}
static float* tp_cand(const b200_ctx* c, uint8_t* base) { return reinterpret_cast<float*>(tp_flags(c, base) + 64); }

extern "C" int b200_ctx_tp_handle(b200_ctx* c, void* handle_out64) {
    if (!c || !handle_out64) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_tp_handle: null argument");
    CU(cudaSetDevice(c->par.device));
    if (!c->tp_region) {
        CU_ALLOC(cudaMalloc((void**)&c->tp_region, tp_region_bytes(c)));
        CU(cudaMemset(c->tp_region, 0, tp_region_bytes(c)));
        CU(cudaDeviceSynchronize());
        c->tp_peer[c->par.rank] = c->tp_region;
        c->tp_peer_set[c->par.rank] = true;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->tp_region));
    memcpy(handle_out64, &h, 64);
    return B200_OK;
}

extern "C" int b200_ctx_tp_set_peer(b200_ctx* c, int peer_rank, const void* handle64) {
    if (!c || !handle64) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_tp_set_peer: null argument");
    if (peer_rank < 0 || peer_rank >= c->par.world_size) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_tp_set_peer: bad rank");
    if (peer_rank == c->par.rank) return B200_OK;
    CU(cudaSetDevice(c->par.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->tp_peer[peer_rank] = (uint8_t*)p;
    c->tp_peer_set[peer_rank] = true;
    return B200_OK;
}

extern "C" void b200_ctx_destroy(b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->par.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->gemv_graph) cudaGraphExecDestroy(c->gemv_graph);
    for (auto& kv : c->tensors)
        if (!kv.second.arena) cudaFree(kv.second.d);
    cudaFree(c->arena);
    for (Slot& s : c->slots) {
        for (int m = 0; m < MODE_COUNT; m++)
            if (s.graph[m]) cudaGraphExecDestroy(s.graph[m]);
        cudaFree(s.d_state);
        cudaFree(s.d_generated);
        cudaFree(s.kv);
        cudaFree(s.kv8);
        cudaFree(s.kv_scale);
        cudaFree(s.d_phases);
        cudaFree(s.d_phases2);
    }
    for (int r = 0; r < c->par.world_size && r < kMmaMaxPeers; r++)
        if (r != c->par.rank && c->tp_peer[r]) cudaIpcCloseMemHandle(c->tp_peer[r]);
    cudaFree(c->tp_region);
    cudaFree(c->mega_bar);
    cudaFree(c->mega_cand_val);
    cudaFree(c->mega_cand_idx);
    cudaFree(c->mega_attn_part);
    cudaFree(c->d_tmaps);
    cudaFree(c->pf_buf);
    cudaFree(c->pf_tok);
    cudaFree(c->pf_rows);
    cudaFree(c->pf_logits);
    cudaFree(c->pf_argmax);
    for (auto& bg : c->batch_graphs)
        if (bg.exec) cudaGraphExecDestroy(bg.exec);
    cudaFree(c->pf_split);
    cudaFree(c->pf_tmaps);
    cudaFree(c->pf_tile_cnt);
    cudaFree(c->attn_q8_part);
    cudaFree(c->pf_k16);
    cudaFree(c->pf_vt16);
    for (uint8_t* p : c->mega_stage) cudaFree(p);
    for (void* p : {(void*)c->xa, (void*)c->xb, (void*)c->qkv, (void*)c->attn, (void*)c->hbuf, (void*)c->logits,
                    (void*)c->attn_part, (void*)c->tickets, (void*)c->moe_sel, (void*)c->moe_wt, (void*)c->taps,
                    (void*)c->rope_freq, c->flush_buf, (void*)c->mma_part, (void*)c->mma_tickets, (void*)c->mma_err})
        cudaFree(p);
    if (c->h_logits) cudaFreeHost(c->h_logits);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->h_token) cudaFreeHost(c->h_token);
    for (int b = 0; b < 2; b++) {
        if (c->h_spec_logits[b]) cudaFreeHost(c->h_spec_logits[b]);
        if (c->spec_ev[b]) cudaEventDestroy(c->spec_ev[b]);
    }
    if (c->h_spec_pick) cudaFreeHost(c->h_spec_pick);
    if (c->h_batch_logits) cudaFreeHost(c->h_batch_logits);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ------------------------------------------------------------------ launches
static void fill_seg(GemvSeg& s, const DevTensor& w, float* out, const DevTensor* bias, int expert_dim) {
    s.w = w.d;
    s.out = out;
    s.bias = (bias && bias->present()) ? bias->f32() : nullptr;
    s.row_bytes = w.row_bytes;
    s.expert_stride = expert_dim ? (long long)(w.nbytes / w.ne[2]) : 0;
    s.type = w.type;
    s.n_rows = (int)w.ne[1];
}

// Tensor-pipe kernel when the launch is eligible (K-quants / Q8_0, shapes that fit), else the CUDA-core kernel.
static bool to_mma_params(b200_ctx* c, const GemvParams& p, MParams& m, MPlan& plan, size_t smem_limit = 0) {
    m = MParams{};
    for (int s = 0; s < p.n_seg; s++) {
        MSeg& d = m.seg[s];
        const GemvSeg& g = p.seg[s];
        d.w = g.w; d.out = g.out; d.bias = g.bias; d.row_bytes = g.row_bytes; d.expert_stride = g.expert_stride;
        d.type = g.type; d.n_rows = g.n_rows;
    }
    m.n_seg = p.n_seg; m.K = p.K; m.x = p.x; m.norm_w = p.norm_w; m.eps = p.eps; m.residual = p.residual;
    m.epi = p.epi == EPI_STORE ? ME_STORE : p.epi == EPI_RESIDUAL ? ME_RESIDUAL : p.epi == EPI_SWIGLU ? ME_SWIGLU : ME_SCALED_ACC;
    m.expert_sel = p.expert_sel; m.expert_wt = p.expert_wt; m.expert_slot = p.expert_slot;
    m.expert_base = p.expert_base; m.expert_count = p.expert_count;
    m.part = c->mma_part; m.tickets = c->mma_tickets; m.err = c->mma_err;
    if (!mma_plan(m, c->n_sm, c->mma_warps, c->mma_stages, smem_limit ? smem_limit : c->smem_optin - 8192, plan)) return false;
    int tiles = 0;
    for (int s = 0; s < m.n_seg; s++) tiles += m.seg[s].n_tiles;
    return tiles <= c->mma_tickets_n;
}

static cudaError_t launch_gemv(b200_ctx* c, GemvParams& p) {
    if (c->use_mma) {
        MParams m;
        MPlan plan;
        if (to_mma_params(c, p, m, plan)) {
            c->mma_launches++;
            return launch_k(c, mma_kernel_for(plan.stages), dim3(plan.grid), dim3(plan.warps * 32), plan.smem, m);
        }
    }
    c->v1_launches++;
    int n_tasks;
    if (p.epi == EPI_SWIGLU) {
        n_tasks = (p.seg[0].n_rows + 1) / 2;
    } else {
        n_tasks = 0;
        for (int s = 0; s < p.n_seg; s++) n_tasks += (p.seg[s].n_rows + kGemvR - 1) / kGemvR;
    }
    int grid = (n_tasks + kGemvWarps - 1) / kGemvWarps;
    grid = std::max(1, std::min(grid, 2 * c->n_sm));
    size_t smem = (size_t)xpad_floats(p.K) * sizeof(float);
    return launch_k(c, gemv_kernel, dim3(grid), dim3(kGemvThreads), smem, p);
}

static cudaError_t launch_attn(b200_ctx* c, const AttnParams& ap, int hd, int G) {
    dim3 grid(ap.n_kv, ap.n_splits);
    if (hd == 128) {
        if (G <= 4) return launch_k(c, attn_decode_kernel<128, 4>, grid, dim3(kAttnThreads), attn_smem_bytes(128, 4, ap.n_splits, ap.G), ap);
        return launch_k(c, attn_decode_kernel<128, 8>, grid, dim3(kAttnThreads), attn_smem_bytes(128, 8, ap.n_splits, ap.G), ap);
    }
    if (G <= 4) return launch_k(c, attn_decode_kernel<64, 4>, grid, dim3(kAttnThreads), attn_smem_bytes(64, 4, ap.n_splits, ap.G), ap);
    return launch_k(c, attn_decode_kernel<64, 8>, grid, dim3(kAttnThreads), attn_smem_bytes(64, 8, ap.n_splits, ap.G), ap);
}

// Enqueue every kernel of one token for `slot` on c->stream (captured into a graph by the caller).
static cudaError_t enqueue_token(b200_ctx* c, int slot_i, Mode mode, bool only_gemv = false) {
    const b200_model_desc& d = c->d;
    Slot& sl = c->slots[slot_i];
    const int H = d.hidden, hd = d.head_dim, nh = d.n_heads, nkv = d.n_kv_heads, G = nh / nkv;
    cudaError_t e;
#define CK(x)                                   \
    do {                                        \
        if ((e = (x)) != cudaSuccess) return e; \
    } while (0)
    const int* pos_ptr = &sl.d_state->pos_cur;
    if (!only_gemv)
        CK(launch_k(c, embed_kernel, dim3(std::max(1, std::min(8, H / 256))), dim3(256), 0, c->token_embd.type,
                    (const uint8_t*)c->token_embd.d, c->token_embd.row_bytes, H, sl.d_state, c->xa, d.vocab));
    if (c->use_taps) CK(cudaMemcpyAsync(c->taps, c->xa, (size_t)H * 4, cudaMemcpyDeviceToDevice, c->stream));
    const size_t kv_layer = (size_t)2 * nkv * d.max_seq_len * hd;
    for (int l = 0; l < d.n_layers; l++) {
        Layer& L = c->layers[l];
        float* kc = sl.kv + (size_t)l * kv_layer;
        float* vc = kc + kv_layer / 2;
        float* q = c->qkv;
        float* k = c->qkv + (size_t)nh * hd;
        float* v = k + (size_t)nkv * hd;
        {   // RMSNorm + QKV (+bias)
            GemvParams p{};
            fill_seg(p.seg[0], L.wq, q, &L.bq, 0);
            fill_seg(p.seg[1], L.wk, k, &L.bk, 0);
            fill_seg(p.seg[2], L.wv, v, &L.bv, 0);
            p.n_seg = 3; p.K = H; p.x = c->xa; p.norm_w = L.attn_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_STORE;
            CK(launch_gemv(c, p));
        }
        if (!only_gemv && c->kv_format == 1) {   // int8 KV cache: RoPE + quantised write, then attention over the dequantised rows
            const size_t rows_layer = (size_t)2 * nkv * d.max_seq_len;
            signed char* k8 = sl.kv8 + (size_t)l * rows_layer * hd;
            signed char* v8 = k8 + rows_layer * hd / 2;
            float* ks = sl.kv_scale + (size_t)l * rows_layer;
            float* vs = ks + rows_layer / 2;
            RopeKvQ8Params rp{};
            rp.q = q; rp.k = k; rp.v = v; rp.k8 = k8; rp.v8 = v8; rp.k_scale = ks; rp.v_scale = vs; rp.freq = c->rope_freq; rp.pos = pos_ptr;
            rp.n_heads = nh; rp.n_kv = nkv; rp.hd = hd; rp.max_seq = d.max_seq_len; rp.neox = d.rope_neox; rp.rope_scale = d.rope_scale;
            const int units = nh + 2 * nkv;
            // (CK is a bare `if`: pick the kernel first, one CK per launch -- an `else CK` would bind to the macro's own `if`)
            CK(launch_k(c, hd == 128 ? rope_kv_q8_kernel<4> : rope_kv_q8_kernel<2>, dim3((units + 3) / 4), dim3(128), 0, rp));
            AttnQ8Params ap{};
            ap.q = q; ap.k8 = k8; ap.v8 = v8; ap.k_scale = ks; ap.v_scale = vs; ap.part = c->attn_q8_part; ap.out = c->attn; ap.pos = pos_ptr;
            ap.n_kv = nkv; ap.G = G; ap.max_seq = d.max_seq_len; ap.n_splits = c->attn_q8_splits; ap.scale = 1.0f / sqrtf((float)hd);
            const dim3 grid(ap.n_splits, nkv);
            CK(launch_k(c, hd == 128 ? attn_q8_split_kernel<128, 4> : attn_q8_split_kernel<64, 4>, grid, dim3(kAttnQ8Warps * 32), 0, ap));
            CK(launch_k(c, attn_q8_merge_kernel, dim3(nh), dim3(hd), 0, ap, hd));
        }
        if (!only_gemv && c->kv_format == 0) {   // RoPE + KV write
            RopeKvParams rp{};
            rp.q = q; rp.k = k; rp.v = v; rp.k_cache = kc; rp.v_cache = vc; rp.freq = c->rope_freq; rp.pos = pos_ptr;
            rp.n_heads = nh; rp.n_kv = nkv; rp.hd = hd; rp.max_seq = d.max_seq_len; rp.neox = d.rope_neox;
            rp.rope_scale = d.rope_scale;
            int work = (nh + nkv) * hd / 2 + nkv * hd;
            CK(launch_k(c, rope_kv_kernel, dim3((work + 255) / 256), dim3(256), 0, rp));
        }
        if (!only_gemv && c->kv_format == 0) {   // GQA decode attention over [0, pos]
            AttnParams ap{};
            ap.q = q; ap.k_cache = kc; ap.v_cache = vc; ap.out = c->attn; ap.part = c->attn_part; ap.tickets = c->tickets;
            ap.pos = pos_ptr; ap.kv_len_fixed = 0; ap.n_kv = nkv; ap.G = G; ap.max_seq = d.max_seq_len;
            ap.n_splits = c->n_splits; ap.scale = 1.0f / sqrtf((float)hd);
            CK(launch_attn(c, ap, hd, G));
        }
        {   // O projection + residual: xb = Wo attn + xa
            GemvParams p{};
            fill_seg(p.seg[0], L.wo, c->xb, nullptr, 0);
            p.n_seg = 1; p.K = nh * hd; p.x = c->attn; p.epi = EPI_RESIDUAL; p.residual = c->xa;
            CK(launch_gemv(c, p));
        }
        if (d.n_experts > 0) {
            RouteParams rt{};
            rt.x = c->xb; rt.norm_w = L.ffn_norm.f32(); rt.eps = d.norm_eps; rt.w_router = L.router.f32();
            rt.hidden = H; rt.n_experts = d.n_experts; rt.top_k = d.n_experts_used; rt.sel = c->moe_sel; rt.wt = c->moe_wt;
            rt.epoch = c->ep ? c->ep_epoch : nullptr;
            if (!only_gemv) CK(launch_k(c, moe_route_kernel, dim3(1), dim3(256), 0, rt));
            for (int s = 0; s < d.n_experts_used; s++) {
                const int e_base = c->ep ? c->par.rank * c->ep_local : 0, e_count = c->ep ? c->ep_local : 0;
                GemvParams p{};
                fill_seg(p.seg[0], L.gate_exps, c->hbuf, nullptr, 1);
                fill_seg(p.seg[1], L.up_exps, c->hbuf, nullptr, 1);
                p.n_seg = 2; p.K = H; p.x = c->xb; p.norm_w = L.ffn_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_SWIGLU;
                p.expert_sel = c->moe_sel; p.expert_wt = c->moe_wt; p.expert_slot = s;
                p.expert_base = e_base; p.expert_count = e_count;
                CK(launch_gemv(c, p));
                GemvParams q2{};
                fill_seg(q2.seg[0], L.down_exps, c->ep ? c->ep_y + (size_t)s * H : c->xa, nullptr, 1);
                q2.n_seg = 1; q2.K = d.expert_ffn; q2.x = c->hbuf; q2.epi = EPI_SCALED_ACC;
                q2.residual = (s == d.n_experts_used - 1) ? c->xb : nullptr;
                q2.expert_sel = c->moe_sel; q2.expert_wt = c->moe_wt; q2.expert_slot = s;
                q2.expert_base = e_base; q2.expert_count = e_count;
                CK(launch_gemv(c, q2));
            }
            if (c->ep && !only_gemv) {   // combine across the GPUs: xa = ((0 + y_0) + y_1 ...) + xb (misc.cuh)
                EpParams ep{};
                ep.sel = c->moe_sel; ep.epoch = c->ep_epoch; ep.top_k = d.n_experts_used; ep.hidden = H;
                ep.rank = c->par.rank; ep.world = c->par.world_size; ep.experts_per_rank = c->ep_local;
                ep.y = c->ep_y;
                for (int r = 0; r < c->par.world_size; r++) {
                    ep.peer_ll[r] = tp_ll(c, c->tp_peer[r], 0, 0);
                    ep.peer_mark[r] = tp_flags(c, c->tp_peer[r]) + 32;
                }
                ep.ll = tp_ll(c, c->tp_region, 0, 0);
                ep.mark = tp_flags(c, c->tp_region) + 32;
                ep.h = c->xb; ep.out = c->xa; ep.err = c->mma_err;
                const int nb = std::max(1, std::min(32, H / 256));
                CK(launch_k(c, ep_push_kernel, dim3(nb), dim3(256), 0, ep));
                CK(launch_k(c, ep_combine_kernel, dim3(nb), dim3(256), 0, ep));
            }
        } else {
            GemvParams p{};
            fill_seg(p.seg[0], L.gate, c->hbuf, nullptr, 0);
            fill_seg(p.seg[1], L.up, c->hbuf, nullptr, 0);
            p.n_seg = 2; p.K = H; p.x = c->xb; p.norm_w = L.ffn_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_SWIGLU;
            CK(launch_gemv(c, p));
            GemvParams q2{};
            fill_seg(q2.seg[0], L.down, c->xa, nullptr, 0);
            q2.n_seg = 1; q2.K = d.ffn; q2.x = c->hbuf; q2.epi = EPI_RESIDUAL; q2.residual = c->xb;
            CK(launch_gemv(c, q2));
        }
        if (c->use_taps)
            CK(cudaMemcpyAsync(c->taps + (size_t)(l + 1) * H, c->xa, (size_t)H * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (mode != MODE_PREFILL) {
        GemvParams p{};
        fill_seg(p.seg[0], c->output.present() ? c->output : c->token_embd, c->logits, nullptr, 0);
        p.n_seg = 1; p.K = H; p.x = c->xa; p.norm_w = c->output_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_STORE;
        CK(launch_gemv(c, p));
        if (mode == MODE_GREEDY && !only_gemv)
            CK(launch_k(c, argmax_kernel, dim3(1), dim3(1024), 0, (const float*)c->logits, d.vocab, sl.d_state,
                        sl.d_generated, kMaxGenerated));
    }
#undef CK
    return cudaSuccess;
}


// ------------------------------------------------------------------ per-token megakernel (mega.cuh)
constexpr int kMegaStages = 3;   // wanted ring depth; a phase whose x + rings do not fit takes fewer (runtime per phase)

static const void* mega_kernel_for(int hd, int G) {
    if (hd == 128) return G <= 4 ? (const void*)mega_decode_kernel<128, 4> : (const void*)mega_decode_kernel<128, 8>;
    return G <= 4 ? (const void*)mega_decode_kernel<64, 4> : (const void*)mega_decode_kernel<64, 8>;
}

static int stream_build(b200_ctx* c);
static int stream2_build(b200_ctx* c, std::vector<std::vector<MegaPhase>>& progs, int max_K);

// Builds the phase program of every slot.  Leaves mega_ok = false (graph path) when a launch is not eligible:
// MoE, taps, a weight type/shape the tensor-pipe GEMV does not take, or a shape that does not fit shared memory.
static int mega_build(b200_ctx* c) {
    const b200_model_desc& d = c->d;
    c->mega_ok = false;
    if (!c->use_mega || !c->use_mma || c->use_taps || d.n_experts > 0) return B200_OK;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->par.device);
    if (!coop) return B200_OK;
    const int H = d.hidden, hd = d.head_dim, nh = d.n_heads, nkv = d.n_kv_heads, G = nh / nkv;
    const size_t lim = c->smem_optin - 8192;
    c->mega_splits = (int)std::max(1, std::min(64, c->n_sm / nkv));
    c->mega_splits = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->mega_splits, (lim - 4096) / ((size_t)G * (hd + 2) * 4)));   // merge scratch must fit
    size_t smem = attn_item_floats(hd, G <= 4 ? 4 : 8, kMmaMaxWarps, c->mega_splits, G) * sizeof(float);
    const size_t kv_layer = (size_t)2 * nkv * d.max_seq_len * hd;
    DevTensor head = c->output.present() ? c->output : c->token_embd;
    const int P = c->par.world_size, R = c->par.rank;
    if (P > 1) {
        if (!c->output.present()) {  // tied head: this rank's rows of the full embedding table
            head.d += (size_t)R * c->vocab_l * head.row_bytes;
            head.nbytes = (size_t)c->vocab_l * head.row_bytes;
        }
        head.ne[1] = (uint64_t)c->vocab_l;
    }
    // tensor parallel: the two row-parallel GEMVs of a layer leave their partial sums in every rank's region
    // (buffer 0: attention output projection, buffer 1: FFN down projection); the next GEMV sums them while staging
    int tp_pending = -1;            // buffer whose sum (+ residual) is the next GEMV's input
    const float* tp_res = nullptr;  // residual to add to that sum
    float* tp_full = nullptr;       // where CTA 0 stores the summed vector (residual of a later phase)
    auto tp_input = [&](MParams& m) {
        if (P <= 1 || tp_pending < 0) return;
        m.xsum = tp_ar(c, c->tp_region, tp_pending, 0);
        m.n_sum = P;
        m.sum_stride = d.hidden;
        m.x_res = tp_res;
        m.x_full_out = tp_full;
        tp_pending = -1;
    };
    auto tp_output = [&](MegaPhase& ph, int buf) {
        if (P <= 1) return;
        MParams& m = ph.gemv;
        m.n_peer = P;
        for (int r = 0; r < P; r++) m.peer_out[r] = tp_ar(c, c->tp_peer[r], buf, R);
        m.epi = ME_STORE;       // the residual is added by the consumer, after the sum
        m.residual = nullptr;
        ph.tp_sync = 1;
    };

    auto gemv_phase = [&](MegaPhase& ph, GemvParams& g) -> bool {
        MParams m;
        MPlan plan;
        const int w = c->mma_warps, st = c->mma_stages;
        c->mma_warps = kMmaMaxWarps;
        c->mma_stages = kMegaStages;
        // (second try without a shared-memory limit: a phase whose x + cp.async rings do not fit the first megakernel -- K = 28672 -- still gets its
        // program entry; mega_fits below decides whether that kernel can run, the streamed kernels have their own plans)
        bool ok = to_mma_params(c, g, m, plan);
        if (!ok) ok = to_mma_params(c, g, m, plan, (size_t)1 << 30);
        c->mma_warps = w;
        c->mma_stages = st;
        if (!ok || plan.warps != kMmaMaxWarps) {
            if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] mega_build: phase not eligible (K %d, %d segments, type %d, rows %d, plan %s, warps %d)\n", g.K, g.n_seg, g.seg[0].type, g.seg[0].n_rows, ok ? "ok" : "failed", plan.warps);
            return false;
        }
        mma_deal(m, c->n_sm);  // every phase runs on the full grid of the megakernel
        ph = MegaPhase{};
        ph.kind = PH_GEMV;
        m.pf_units = env_int("B200_PF_UNITS", 8);
        ph.gemv = m;
        tp_input(ph.gemv);
        smem = std::max(smem, plan.smem);
        return true;
    };

    if (!c->mega_bar) {
        CU_ALLOC(cudaMalloc((void**)&c->mega_bar, sizeof(unsigned int)));
        CU_ALLOC(cudaMalloc((void**)&c->mega_cand_val, c->n_sm * sizeof(float)));
        CU_ALLOC(cudaMalloc((void**)&c->mega_cand_idx, c->n_sm * sizeof(int)));
        CU_ALLOC(cudaMalloc((void**)&c->mega_attn_part, (size_t)nkv * c->mega_splits * G * (hd + 2) * sizeof(float)));
    }
    // single GPU: every activation vector that feeds a GEMV is also kept in its staged form, written by the phase
    // that produces it (gemv_mma.cuh: stage_out32), so the consuming GEMV copies it instead of converting it per CTA
    const bool staged = P == 1 && env_int("B200_STAGED_X", 1) && H % 32 == 0 && (nh * hd) % 32 == 0 && d.ffn % 32 == 0;
    if (staged && !c->mega_stage[0]) {
        const int ks[4] = {H, H, nh * hd, (int)d.ffn};
        for (int i = 0; i < 4; i++) {
            CU_ALLOC(cudaMalloc((void**)&c->mega_stage[i], x_staged_bytes(ks[i]) + 256));
            CU(cudaMemset(c->mega_stage[i], 0, x_staged_bytes(ks[i]) + 256));
        }
    }
    uint8_t* const st_xa = staged ? c->mega_stage[0] : nullptr;
    uint8_t* const st_xb = staged ? c->mega_stage[1] : nullptr;
    uint8_t* const st_attn = staged ? c->mega_stage[2] : nullptr;
    uint8_t* const st_hbuf = staged ? c->mega_stage[3] : nullptr;
    for (size_t si = 0; si < c->slots.size(); si++) {
        Slot& sl = c->slots[si];
        std::vector<MegaPhase> prog;
        for (int l = 0; l < d.n_layers; l++) {
            Layer& L = c->layers[l];
            float* kc = sl.kv + (size_t)l * kv_layer;
            float* vc = kc + kv_layer / 2;
            MegaPhase ph;
            {   // RMSNorm + QKV (+bias) -> raw q | k | v
                GemvParams p{};
                fill_seg(p.seg[0], L.wq, c->qkv, &L.bq, 0);
                fill_seg(p.seg[1], L.wk, c->qkv + (size_t)nh * hd, &L.bk, 0);
                fill_seg(p.seg[2], L.wv, c->qkv + (size_t)(nh + nkv) * hd, &L.bv, 0);
                p.n_seg = 3; p.K = H; p.x = c->xa; p.norm_w = L.attn_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_STORE;
                tp_res = c->xb; tp_full = c->xa;   // TP, layers > 0: xa = sum(down partials) + xb
                if (!gemv_phase(ph, p)) return B200_OK;
                if (l > 0) ph.gemv.x_staged = st_xa;   // layer 0 reads the embedding row (f32)
                prog.push_back(ph);
            }
            {   // RoPE + KV write + GQA decode attention
                ph = MegaPhase{};
                ph.kind = PH_ATTN;
                AttnParams& ap = ph.attn;
                ap.q = nullptr; ap.k_cache = kc; ap.v_cache = vc; ap.out = c->attn; ap.part = c->mega_attn_part; ap.tickets = c->tickets;
                ap.pos = &sl.d_state->pos_cur; ap.kv_len_fixed = 0; ap.n_kv = nkv; ap.G = G; ap.max_seq = d.max_seq_len;
                ap.n_splits = c->mega_splits; ap.scale = 1.0f / sqrtf((float)hd);
                ap.min_chunk = env_int("B200_ATTN_MIN_CHUNK", 0);
                ap.qkv_raw = c->qkv; ap.freq = c->rope_freq; ap.rope_scale = d.rope_scale; ap.neox = d.rope_neox; ap.n_heads = nh;
                ap.stage_out = st_attn; ap.stage_K = nh * hd;
                prog.push_back(ph);
            }
            {   // O projection + residual: xb = Wo attn + xa
                GemvParams p{};
                fill_seg(p.seg[0], L.wo, c->xb, nullptr, 0);
                p.n_seg = 1; p.K = nh * hd; p.x = c->attn; p.epi = EPI_RESIDUAL; p.residual = c->xa;
                if (!gemv_phase(ph, p)) return B200_OK;
                tp_output(ph, 0);
                ph.gemv.x_staged = st_attn;
                ph.gemv.stage_out = st_xb; ph.gemv.stage_w = L.ffn_norm.f32(); ph.gemv.stage_K = H;
                prog.push_back(ph);
                if (P > 1) { tp_pending = 0; tp_res = c->xa; tp_full = c->xb; }   // xb = sum(O partials) + xa
            }
            {   // RMSNorm + gate | up + SwiGLU
                GemvParams p{};
                fill_seg(p.seg[0], L.gate, c->hbuf, nullptr, 0);
                fill_seg(p.seg[1], L.up, c->hbuf, nullptr, 0);
                p.n_seg = 2; p.K = H; p.x = c->xb; p.norm_w = L.ffn_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_SWIGLU;
                if (!gemv_phase(ph, p)) return B200_OK;
                ph.gemv.x_staged = st_xb;
                ph.gemv.stage_out = st_hbuf; ph.gemv.stage_w = nullptr; ph.gemv.stage_K = (int)d.ffn;
                prog.push_back(ph);
            }
            {   // down + residual: xa = Wd act + xb
                GemvParams p{};
                fill_seg(p.seg[0], L.down, c->xa, nullptr, 0);
                p.n_seg = 1; p.K = d.ffn; p.x = c->hbuf; p.epi = EPI_RESIDUAL; p.residual = c->xb;
                if (!gemv_phase(ph, p)) return B200_OK;
                tp_output(ph, 1);
                ph.gemv.x_staged = st_hbuf;
                ph.gemv.stage_out = st_xa; ph.gemv.stage_K = H;
                ph.gemv.stage_w = (l + 1 < d.n_layers) ? c->layers[l + 1].attn_norm.f32() : c->output_norm.f32();
                prog.push_back(ph);
                if (P > 1) { tp_pending = 1; tp_res = c->xb; tp_full = c->xa; }   // xa = sum(down partials) + xb
            }
        }
        {   // final RMSNorm + vocab head
            MegaPhase ph;
            GemvParams p{};
            fill_seg(p.seg[0], head, c->logits, nullptr, 0);
            p.n_seg = 1; p.K = H; p.x = c->xa; p.norm_w = c->output_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_STORE;
            if (!gemv_phase(ph, p)) return B200_OK;
            ph.gemv.x_staged = st_xa;
            prog.push_back(ph);
        }
        if (sl.d_phases) cudaFree(sl.d_phases);
        CU_ALLOC(cudaMalloc((void**)&sl.d_phases, prog.size() * sizeof(MegaPhase)));
        CU(cudaMemcpy(sl.d_phases, prog.data(), prog.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice));
        c->mega_phases = (int)prog.size();
    }
    // The first megakernel keeps x AND a cp.async ring per warp in shared memory: K = 28672 (Llama-3-70B's FFN) does not fit.  The streamed
    // kernels only need the phase program built above, so they are tried either way; mega_ok then means "some per-token megakernel runs".
    bool mega_fits = smem <= lim;
    c->mega_smem = smem;
    if (mega_fits) {
        const void* kern = mega_kernel_for(hd, G);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMmaMaxWarps * 32, smem));
        mega_fits = per_sm >= 1;
    }
    c->mega_ok = true;   // (stream_build / stream2_build read the phase programs; they do not look at this flag)
    const int rc = stream_build(c);
    if (rc != B200_OK) return rc;
    if (!mega_fits) c->mega_ok = c->stream_ok || c->stream2_ok;
    if (!mega_fits && env_int("B200_LOG", 0)) fprintf(stderr, "[b200] mega_build: first megakernel does not fit (%zu bytes of shared memory), streamed kernel %s\n", smem, c->mega_ok ? "runs" : "not eligible either");
    return B200_OK;
}


// ------------------------------------------------------------------ streamed megakernel (stream.cuh)
static const void* stream_kernel_for(int hd, int G, bool tp) {
    if (tp) {
        if (hd == 128) return G <= 4 ? (const void*)stream_decode_kernel<128, 4, true> : (const void*)stream_decode_kernel<128, 8, true>;
        return G <= 4 ? (const void*)stream_decode_kernel<64, 4, true> : (const void*)stream_decode_kernel<64, 8, true>;
    }
    if (hd == 128) return G <= 4 ? (const void*)stream_decode_kernel<128, 4, false> : (const void*)stream_decode_kernel<128, 8, false>;
    return G <= 4 ? (const void*)stream_decode_kernel<64, 4, false> : (const void*)stream_decode_kernel<64, 8, false>;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Upgrades the phase programs built by mega_build to the streamed kernel when every weight matrix can be described
// by a TMA tensor map (rows that are 16-byte multiples, 16-byte aligned base) and every phase uses one entry shape.
// Leaves stream_ok = false (first megakernel) otherwise.
static int stream_build(b200_ctx* c) {
    const b200_model_desc& d = c->d;
    c->stream_ok = false;
    if (!c->use_stream || (c->par.world_size > 1 && !env_int("B200_STREAM_TP", 1))) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: not eligible (check 1, line %d)\n", __LINE__); return B200_OK; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: no cuTensorMapEncodeTiled entry point\n");
        return B200_OK;
    }
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    const int hd = d.head_dim, G = d.n_heads / d.n_kv_heads;
    std::vector<CUtensorMap> maps;
    std::map<const void*, int> map_of;   // weight base -> index in maps
    std::vector<std::vector<MegaPhase>> progs(c->slots.size());
    int max_K = d.hidden;
    for (size_t si = 0; si < c->slots.size(); si++) {
        std::vector<MegaPhase>& prog = progs[si];
        prog.resize(c->mega_phases);
        CU(cudaMemcpy(prog.data(), c->slots[si].d_phases, prog.size() * sizeof(MegaPhase), cudaMemcpyDeviceToHost));
        for (size_t pi = 0; pi < prog.size(); pi++) {
            if (prog[pi].kind != PH_GEMV) continue;
            MParams& m = prog[pi].gemv;
            if (m.expert_sel || m.K % kMmaChunk) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: not eligible (check 2, line %d)\n", __LINE__); return B200_OK; }
            int C = 1 << 30;
            for (int s = 0; s < m.n_seg; s++) {
                if (!mma_type_ok(m.seg[s].type)) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: not eligible (type %d)\n", m.seg[s].type); return B200_OK; }
                C = std::min(C, stream_pref_chunks(m.seg[s].type));
            }
            {   // short phases: a CTA that owns only a handful of entries cannot spread them over its warps (O projection:
                // one tile = 8 two-chunk entries for 7 warps, one warp does twice the work of the others) -> one-chunk entries
                int tiles_all = 0;
                for (int s = 0; s < (m.epi == ME_SWIGLU ? 1 : m.n_seg); s++) tiles_all += m.seg[s].n_tiles;
                const int per_cta = (tiles_all + c->n_sm - 1) / c->n_sm;
                const int entries = per_cta * (m.epi == ME_SWIGLU ? 2 : 1) * ((m.chunks + C - 1) / C);
                if (C > 1 && entries < env_int("B200_STREAM_MIN_ENTRIES", 0)) C = 1;
                if (C > 1 && c->use_stream2 && c->par.world_size == 1) {
                    // stream2.cuh deals work as jobs = (entry, chunk): two-chunk entries halve the TMA operations the single producer
                    // thread has to issue without coarsening the deal; every entry must then have both chunks
                    if ((m.chunks & 1) || env_int("B200_S2_C", 0) == 1) C = 1;
                }
            }
            for (int s = 0; s < m.n_seg; s++) {
                MSeg& sg = m.seg[s];
                if ((sg.row_bytes & 15) || ((uintptr_t)sg.w & 15)) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: not eligible (row bytes %lld)\n", sg.row_bytes); return B200_OK; }
                sg.s_pitch = stream_pitch(sg.type, C);
                sg.s_elem = 4;
                if (sg.s_pitch * kMmaRows > kStreamSlotBytes || sg.s_pitch / 4 > 256) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: not eligible (pitch %d)\n", sg.s_pitch); return B200_OK; }
                auto it = map_of.find(sg.w);
                if (it == map_of.end()) {
                    CUtensorMap tm;
                    const cuuint64_t dims[2] = {(cuuint64_t)(sg.row_bytes / 4), (cuuint64_t)sg.n_rows};
                    const cuuint64_t strides[1] = {(cuuint64_t)sg.row_bytes};
                    const cuuint32_t box[2] = {(cuuint32_t)(sg.s_pitch / 4), (cuuint32_t)kMmaRows};
                    const cuuint32_t estr[2] = {1, 1};
                    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)sg.w, dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: cuTensorMapEncodeTiled failed (%d) for type %d rows %d row_bytes %lld\n", (int)r, sg.type, sg.n_rows, sg.row_bytes); return B200_OK; }
                    it = map_of.emplace(sg.w, (int)maps.size()).first;
                    maps.push_back(tm);
                }
                sg.tmap = (const void*)(uintptr_t)(it->second + 1);   // index + 1 for now; device address below
            }
            m.s_C = C;
            m.s_ept = (m.chunks + C - 1) / C;
            m.s_parts = m.epi == ME_SWIGLU ? 2 : 1;
            int tiles = 0;
            for (int s = 0; s < (m.epi == ME_SWIGLU ? 1 : m.n_seg); s++) tiles += m.seg[s].n_tiles;
            m.s_tiles = tiles;
            m.s_ncta = std::min(c->n_sm, tiles);
            m.s_cbase = tiles / m.s_ncta;
            m.s_crem = tiles % m.s_ncta;
            m.s_rot = env_int("B200_STREAM_ROT", 1) ? (int)((pi * 37) % (size_t)c->n_sm) : 0;
            max_K = std::max(max_K, m.K);
        }
    }
    const void* kern = stream_kernel_for(hd, G, c->par.world_size > 1);
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kern));
    size_t x_region = std::max(x_smem_bytes(max_K) + 16, attn_item_floats(hd, G <= 4 ? 4 : 8, kSW, c->mega_splits, G) * sizeof(float));
    x_region = (x_region + 127) & ~(size_t)127;
    const size_t avail = c->smem_optin - fa.sharedSizeBytes;
    // whether the FIRST streamed kernel can run (the second one has its own shared-memory plan: stream2_build, tried either way)
    bool s1_ok = avail >= x_region + 3 * (size_t)kStreamSlotBytes;
    int slots = 0;
    size_t smem = 0;
    if (s1_ok) {
        slots = (int)std::min<size_t>(kStreamMaxSlots, (avail - x_region) / kStreamSlotBytes);
        slots = std::min(slots, std::max(2, env_int("B200_STREAM_SLOTS", kStreamMaxSlots)));
        s1_ok = slots > kSW;   // the parity protocol needs more slots than consumer warps
    }
    if (s1_ok) {
        smem = x_region + (size_t)slots * kStreamSlotBytes;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStreamThreads, smem));
        s1_ok = per_sm >= 1;
    }
    if (!s1_ok && env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: first streamed kernel not eligible (x region %zu of %zu bytes, %d slots)\n", x_region, avail, slots);
    if (c->d_tmaps) cudaFree(c->d_tmaps);
    c->d_tmaps = nullptr;
    CU_ALLOC(cudaMalloc(&c->d_tmaps, maps.size() * sizeof(CUtensorMap)));
    CU(cudaMemcpy(c->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    for (size_t si = 0; si < c->slots.size(); si++) {
        for (MegaPhase& ph : progs[si]) {
            if (ph.kind != PH_GEMV) continue;
            for (int s = 0; s < ph.gemv.n_seg; s++)
                ph.gemv.seg[s].tmap = (const CUtensorMap*)c->d_tmaps + ((int)(uintptr_t)ph.gemv.seg[s].tmap - 1);
        }
        CU(cudaMemcpy(c->slots[si].d_phases, progs[si].data(), progs[si].size() * sizeof(MegaPhase), cudaMemcpyHostToDevice));
    }
    c->stream_smem = smem;
    c->stream_ring_off = (int)x_region;
    c->stream_slots = slots;
    c->stream_ok = s1_ok;
    if (s1_ok && env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream_build: ok, %d slots, ring_off %d, smem %zu, %zu tensor maps\n", slots, (int)x_region, smem, maps.size());
    return stream2_build(c, progs, max_K);
}

static const void* stream2_kernel_for(int hd, int G) {
    if (hd == 128) return G <= 4 ? (const void*)stream2_decode_kernel<128, 4> : (const void*)stream2_decode_kernel<128, 8>;
    return G <= 4 ? (const void*)stream2_decode_kernel<64, 4> : (const void*)stream2_decode_kernel<64, 8>;
}

// Upgrades the (stream-eligible) phase programs to the second streamed megakernel: single GPU, dense.  The program gets an
// EMBED phase in front (CTA 0: pick of the previous token in greedy mode, embedding row, its staged form for layer 0's
// QKV), every GEMV reads a staged input, the vocab head collects argmax candidates.  Leaves stream2_ok = false otherwise.
static int stream2_build(b200_ctx* c, std::vector<std::vector<MegaPhase>>& progs, int max_K) {
    const b200_model_desc& d = c->d;
    c->stream2_ok = false;
    if (!c->use_stream2 || (c->par.world_size > 1 && !env_int("B200_STREAM2_TP", 1))) return B200_OK;
    const int P = c->par.world_size;
    const bool use_ll = P > 1 && env_int("B200_TP_LL", 1);       // all-reduce inside the row-parallel GEMV's phase (packets); 0: flag exchange + REDUCE phase
    const bool tp_fold = P > 1 && env_int("B200_TP_FOLD", 0);   // (measured slower) REDUCE folded into the consuming GEMV's staging
    const int hd = d.head_dim, G = d.n_heads / d.n_kv_heads, gmax = G <= 4 ? 4 : 8;
    if (!c->mega_stage[0]) {   // tensor parallel: mega_build keeps no staged vectors (its kernels sum the partials while staging); this one does
        if (P == 1 || d.hidden % 32 || (d.n_heads * hd) % 32 || d.ffn % 32) return B200_OK;
        const int ks[4] = {d.hidden, d.hidden, d.n_heads * hd, (int)d.ffn};
        for (int i = 0; i < 4; i++) {
            CU_ALLOC(cudaMalloc((void**)&c->mega_stage[i], x_staged_bytes(ks[i]) + 256));
            CU(cudaMemset(c->mega_stage[i], 0, x_staged_bytes(ks[i]) + 256));
        }
    }
    const void* kern = stream2_kernel_for(hd, G);
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kern));
    // KV splits per kv head: the ticket merge stages [splits][G][hd + 2] floats in the x region.  Few kv heads per GPU (tensor parallel,
    // G = 8: 37 splits x 8 x 130 floats = 154 KB) would leave no room for the ring: the splits are capped so that the scratch stays within
    // the larger of the GEMV input and 64 KB (the descriptors of this kernel carry their own n_splits; partials are indexed with it)
    int s2_splits = c->mega_splits;
    {
        const size_t budget = std::max(x_smem_bytes(max_K) + 16, (size_t)64 << 10);
        while (s2_splits > 1 && attn2_smem_floats(hd, gmax, kS2Cons, s2_splits, G) * sizeof(float) > budget) s2_splits--;
    }
    size_t x_region = std::max(x_smem_bytes(max_K) + 16, attn2_smem_floats(hd, gmax, kS2Cons, s2_splits, G) * sizeof(float));
    x_region = (x_region + 127) & ~(size_t)127;
    const size_t xr_off = kS2ZeroBytes;
    const size_t tpart_off = xr_off + x_region;
    const size_t desc_off = tpart_off + (size_t)kS2TileSlots * kS2Cons * 2 * 32 * sizeof(float);
    const size_t ring_off = (desc_off + 2 * sizeof(MegaPhase) + 127) & ~(size_t)127;
    const size_t avail = c->smem_optin - fa.sharedSizeBytes;
    if (avail < ring_off + (size_t)kS2Cons * kS2SlotBytes) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream2_build: not eligible (shared memory)\n"); return B200_OK; }
    int slots = (int)std::min<size_t>(kS2MaxSlots, (avail - ring_off) / kS2SlotBytes);
    slots = std::min(slots, std::max(2, env_int("B200_STREAM_SLOTS", kS2MaxSlots)));
    slots = slots / kS2Cons * kS2Cons;       // a multiple of 14: a slot always serves the same consumer warps (no parity aliasing)
    if (slots < kS2Cons) return B200_OK;
    const size_t smem = ring_off + (size_t)slots * kS2SlotBytes;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kS2Threads, smem));
    if (per_sm < 1) { if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream2_build: not eligible (occupancy)\n"); return B200_OK; }
    if (!c->s2_ll) {
        CU_ALLOC(cudaMalloc((void**)&c->s2_ll, (size_t)c->n_sm * 64 * sizeof(uint2)));
        CU(cudaMemset(c->s2_ll, 0, (size_t)c->n_sm * 64 * sizeof(uint2)));
        CU_ALLOC(cudaMalloc((void**)&c->s2_cand_val, (size_t)c->n_sm * kS2Cons * sizeof(float)));
        CU_ALLOC(cudaMalloc((void**)&c->s2_cand_idx, (size_t)c->n_sm * kS2Cons * sizeof(int)));
        CU(cudaMemset(c->s2_cand_idx, 0xff, (size_t)c->n_sm * kS2Cons * sizeof(int)));
        CU(cudaMemset(c->s2_cand_val, 0, (size_t)c->n_sm * kS2Cons * sizeof(float)));
    }
    const int min_chunk2 = env_int("B200_ATTN_MIN_CHUNK", 32);   // measured: 11.5 us per layer at kv_len 144 with 32, 16.7 with 256
    for (size_t si = 0; si < c->slots.size(); si++) {
        std::vector<MegaPhase> prog2;
        MegaPhase em{};
        em.kind = PH_EMBED;
        em.gemv.stage_out = c->mega_stage[0];
        em.gemv.stage_w = c->layers[0].attn_norm.f32();
        em.gemv.stage_K = d.hidden;
        prog2.push_back(em);
        for (size_t pi = 0; pi < progs[si].size(); pi++) {
            MegaPhase ph = progs[si][pi];
            const int li = (int)(pi / 5), k5 = (int)(pi % 5);          // mega_build's program: 5 phases per layer (QKV, ATTN, O, GATE/UP, DOWN), then the head
            const bool is_head = pi + 1 == progs[si].size();
            MegaPhase red{};                                            // tensor parallel: the REDUCE phase that follows a row-parallel GEMV
            bool have_red = false;
            if (ph.kind == PH_GEMV) {
                MParams& m = ph.gemv;
                if (P == 1) {
                    m.x_staged = pi == 0 ? c->mega_stage[0] : m.x_staged;   // layer 0's QKV reads the staged embedding row
                } else {
                    // the staged-vector wiring mega_build does for one GPU, plus the all-reduce: the ranks' partial vectors are summed (in
                    // rank order, + residual) by a REDUCE phase that also writes the staged form the next GEMV copies
                    const float* next_norm = is_head ? nullptr : (li + 1 < d.n_layers ? c->layers[li + 1].attn_norm.f32() : c->output_norm.f32());
                    // B200_TP_FOLD=1 (default): no REDUCE phase -- the GEMV that consumes the sum keeps mega_build's xsum / x_res / x_full_out
                    // wiring and its consumer warps finish the all-reduce while they stage x in shared memory (stream2.cuh: s2_gemv_cta)
                    const bool fold = tp_fold && !use_ll && m.n_sum > 0;
                    if (!fold) { m.xsum = nullptr; m.n_sum = 0; m.x_res = nullptr; m.x_full_out = nullptr; }
                    if (fold) {
                        if (!m.norm_w || m.K != d.hidden || (k5 != 0 && k5 != 3 && !is_head)) return B200_OK;
                        m.x_staged = nullptr;
                        if (k5 == 3) { m.stage_out = c->mega_stage[3]; m.stage_w = nullptr; m.stage_K = (int)d.ffn; }
                    } else if (use_ll && (k5 == 2 || k5 == 4)) {
                        // row-parallel GEMV with the all-reduce inside its phase (stream2.cuh: ll_red): packets to every rank, polled by the
                        // consumer warps; buffer 0 / xb = sum + xa staged with ffn_norm (O), buffer 1 / xa = sum + xb staged with the next norm (down)
                        const int buf = k5 == 2 ? 0 : 1;
                        if (!ph.tp_sync || m.n_peer != P) return B200_OK;
                        m.x_staged = c->mega_stage[k5 == 2 ? 2 : 3];
                        m.ll_red = 1;
                        ph.tp_sync = 0;
                        for (int r = 0; r < P; r++) m.peer_out[r] = reinterpret_cast<float*>(tp_ll(c, c->tp_peer[r], buf, c->par.rank));
                        m.xsum = reinterpret_cast<const float*>(tp_ll(c, c->tp_region, buf, 0));
                        m.n_sum = P; m.sum_stride = d.hidden;
                        m.x_res = k5 == 2 ? c->xa : c->xb;
                        m.x_full_out = k5 == 2 ? c->xb : c->xa;
                        m.stage_out = c->mega_stage[k5 == 2 ? 1 : 0];
                        m.stage_w = k5 == 2 ? c->layers[li].ffn_norm.f32() : next_norm;
                        m.stage_K = d.hidden;
                    } else if (use_ll && k5 == 3) {
                        m.x_staged = c->mega_stage[1];
                        m.stage_out = c->mega_stage[3]; m.stage_w = nullptr; m.stage_K = (int)d.ffn;
                    } else if (use_ll) {   // QKV, vocab head
                        m.x_staged = c->mega_stage[0];
                    } else if (tp_fold && (k5 == 2 || k5 == 4)) {   // row-parallel: partial vectors to every rank, nothing staged
                        m.x_staged = c->mega_stage[k5 == 2 ? 2 : 3];
                        m.stage_out = nullptr;
                        if (!ph.tp_sync || m.n_peer != P) return B200_OK;
                    } else if (is_head || k5 == 0) {
                        m.x_staged = c->mega_stage[0];
                    } else if (k5 == 2) {            // O projection: partial vectors -> buffer 0 of every rank; REDUCE: xb = sum + xa, staged with ffn_norm
                        m.x_staged = c->mega_stage[2];
                        m.stage_out = nullptr;
                        red.gemv.xsum = tp_ar(c, c->tp_region, 0, 0); red.gemv.x_res = c->xa; red.gemv.x_full_out = c->xb;
                        red.gemv.stage_out = c->mega_stage[1]; red.gemv.stage_w = c->layers[li].ffn_norm.f32();
                        have_red = true;
                    } else if (k5 == 3) {            // gate | up + SwiGLU -> staged hbuf (this rank's slice of the FFN)
                        m.x_staged = c->mega_stage[1];
                        m.stage_out = c->mega_stage[3]; m.stage_w = nullptr; m.stage_K = (int)d.ffn;
                    } else if (k5 == 4) {            // down projection: partial vectors -> buffer 1; REDUCE: xa = sum + xb, staged with the next norm
                        m.x_staged = c->mega_stage[3];
                        m.stage_out = nullptr;
                        red.gemv.xsum = tp_ar(c, c->tp_region, 1, 0); red.gemv.x_res = c->xb; red.gemv.x_full_out = c->xa;
                        red.gemv.stage_out = c->mega_stage[0]; red.gemv.stage_w = next_norm;
                        have_red = true;
                    }
                    if (have_red) {
                        red.kind = PH_REDUCE;
                        red.gemv.K = d.hidden; red.gemv.stage_K = d.hidden;
                        red.gemv.n_sum = P; red.gemv.sum_stride = d.hidden;
                        if (!ph.tp_sync || m.n_peer != P) return B200_OK;   // (mega_build marks the row-parallel phases)
                    }
                }
                if (!m.x_staged && m.n_sum == 0) return B200_OK;
                m.s_E = m.s_tiles * m.s_parts * m.s_ept;
                m.cand = is_head ? 1 : 0;
                {   // jobs per entry (stream2.cuh): the coarsest split whose last round of the 14 consumer warps is (nearly) as full as the best
                    const double per = (double)m.s_E / c->n_sm;
                    auto waste = [&](int n) { return std::ceil(per * n / kS2Cons) * kS2Cons / std::max(per * n, 1e-9); };
                    const int cand_J[3] = {1, m.s_C == 2 ? 2 : 1, m.s_C == 2 ? 2 : 1}, cand_R[3] = {1, 1, 2};
                    double best = 1e9;
                    const int n_cand = kS2Cons == 8 ? 2 : 3;   // (8 fat consumers: whole units only)
                    for (int k = 0; k < n_cand; k++) best = std::min(best, waste(cand_J[k] * cand_R[k]));
                    int pick = n_cand - 1;
                    for (int k = 0; k < n_cand; k++)
                        if (waste(cand_J[k] * cand_R[k]) <= best + 0.09) { pick = k; break; }   // (coarser jobs amortise the per-job overhead: measured)
                    const int force = env_int("B200_S2_JOBS", 0);   // 1, 2 or 4 jobs per entry
                    if (force == 1) pick = 0; else if (force == 2) pick = 1; else if (force == 4 && n_cand == 3) pick = 2;
                    m.s_J = cand_J[pick]; m.s_R = cand_R[pick];
                    m.s_jsh = (m.s_J == 2 ? 1 : 0) + (m.s_R == 2 ? 1 : 0);
                    if (env_int("B200_LOG", 0) && si == 0 && pi < 6) fprintf(stderr, "[b200] stream2 phase %zu: E %d (%.1f per CTA), C %d, jobs per entry %d x %d, waste %.3f\n", pi, m.s_E, per, m.s_C, m.s_J, m.s_R, waste(m.s_J * m.s_R));
                }
                m.x_bytes = m.x_staged ? (int)x_staged_bytes(m.K) : 0;
            } else {
                ph.attn.min_chunk = min_chunk2;
                ph.attn.n_splits = s2_splits;
                if (P > 1) { ph.attn.stage_out = c->mega_stage[2]; ph.attn.stage_K = d.n_heads * hd; }
            }
            prog2.push_back(ph);
            if (have_red) prog2.push_back(red);
        }
        Slot& sl = c->slots[si];
        if (sl.d_phases2) cudaFree(sl.d_phases2);
        sl.d_phases2 = nullptr;
        CU_ALLOC(cudaMalloc((void**)&sl.d_phases2, prog2.size() * sizeof(MegaPhase)));
        CU(cudaMemcpy(sl.d_phases2, prog2.data(), prog2.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice));
        c->s2_phases = (int)prog2.size();
    }
    c->s2_xr_off = (int)xr_off; c->s2_tpart_off = (int)tpart_off; c->s2_desc_off = (int)desc_off; c->s2_ring_off = (int)ring_off;
    c->s2_slots = slots;
    c->s2_smem = smem;
    c->stream2_ok = true;
    if (env_int("B200_LOG", 0)) fprintf(stderr, "[b200] stream2_build: ok, %d slots, x region %zu, ring_off %zu, smem %zu, %d phases, %d regs, %d KV splits\n", slots, x_region, ring_off, smem, c->s2_phases, fa.numRegs, s2_splits);
    return B200_OK;
}

// One cooperative launch: n_tokens tokens of `slot` (MEGA_GREEDY feeds its own argmax back in).
static int mega_launch(b200_ctx* c, int slot_i, int mode, int n_tokens) {
    Slot& sl = c->slots[slot_i];
    const b200_model_desc& d = c->d;
    MegaParams mp{};
    mp.phases = sl.d_phases; mp.n_phases = c->mega_phases; mp.mode = mode; mp.n_tokens = n_tokens;
    mp.bar = c->mega_bar; mp.err = c->mma_err;
    mp.embd_type = c->token_embd.type; mp.embd = c->token_embd.d; mp.embd_row_bytes = c->token_embd.row_bytes;
    mp.hidden = d.hidden; mp.vocab = d.vocab; mp.h = c->xa;
    mp.st = sl.d_state; mp.logits = c->logits; mp.cand_val = c->mega_cand_val; mp.cand_idx = c->mega_cand_idx;
    mp.generated = sl.d_generated; mp.max_generated = kMaxGenerated;
    mp.hd = d.head_dim; mp.G = d.n_heads / d.n_kv_heads;
    mp.dbg = c->mega_dbg;
    mp.early = env_int("B200_MEGA_EARLY", 1);
    mp.tp_size = c->par.world_size; mp.tp_rank = c->par.rank; mp.vocab_local = c->vocab_l;
    if (c->par.world_size > 1) {
        mp.tp_flags = tp_flags(c, c->tp_region);
        mp.tp_cand = tp_cand(c, c->tp_region);
        for (int r = 0; r < c->par.world_size; r++) {
            mp.tp_peer_flags[r] = tp_flags(c, c->tp_peer[r]);
            mp.tp_peer_cand[r] = tp_cand(c, c->tp_peer[r]);
        }
        mp.tp_epoch0 = c->tp_epoch;
        const int per_token = 2 * d.n_layers + (mode == MEGA_GREEDY ? 1 : 0);
        c->tp_epoch += (unsigned int)(per_token * n_tokens) + (c->stream2_ok ? 1u : 0u);   // (stream2: the last token's pick is one more exchange)
    }
    CU(cudaMemsetAsync(c->mega_bar, 0, sizeof(unsigned int), c->stream));
    if (c->stream2_ok) {
        Stream2Params sp{};
        sp.mp = mp;
        sp.mp.phases = sl.d_phases2;
        sp.mp.n_phases = c->s2_phases;
        sp.xr_off = c->s2_xr_off; sp.tpart_off = c->s2_tpart_off; sp.desc_off = c->s2_desc_off; sp.ring_off = c->s2_ring_off;
        sp.n_slots = c->s2_slots;
        sp.no_load = env_int("B200_STREAM_NOLOAD", 0);
        sp.ll = c->s2_ll;
        sp.tp_per_token = 2 * d.n_layers + (mode == MEGA_GREEDY ? 1 : 0);
        sp.epoch0 = c->s2_epoch;
        c->s2_epoch += (unsigned int)(n_tokens * c->s2_phases + 2);
        sp.cand_val = c->s2_cand_val; sp.cand_idx = c->s2_cand_idx;
        void* sargs[] = {&sp};
        CU(cudaLaunchCooperativeKernel(stream2_kernel_for(d.head_dim, d.n_heads / d.n_kv_heads), dim3(c->n_sm), dim3(kS2Threads), sargs, c->s2_smem,
                                       c->stream));
    } else if (c->stream_ok) {
        StreamParams sp{};
        sp.mp = mp;
        sp.ring_off = c->stream_ring_off;
        sp.n_slots = c->stream_slots;
        sp.no_load = env_int("B200_STREAM_NOLOAD", 0);
        void* sargs[] = {&sp};
        CU(cudaLaunchCooperativeKernel(stream_kernel_for(d.head_dim, d.n_heads / d.n_kv_heads, c->par.world_size > 1), dim3(c->n_sm), dim3(kStreamThreads), sargs,
                                       c->stream_smem, c->stream));
    } else {
        void* args[] = {&mp};
        CU(cudaLaunchCooperativeKernel(mega_kernel_for(d.head_dim, d.n_heads / d.n_kv_heads), dim3(c->n_sm), dim3(kMmaMaxWarps * 32), args,
                                       c->mega_smem, c->stream));
    }
    c->launches += 1;
    c->mega_launches += 1;
    return B200_OK;
}


// ------------------------------------------------------------------ GEMM prefill
// A prompt of n >= prefill_gemm_min tokens is processed prefill_chunk() tokens at a time: every weight matrix is one
// tcgen05 dequant-GEMM per chunk (gemm_umma.cuh) instead of one GEMV per token (what the reference does:
// src/model/llama.rs:327-345), RoPE / KV write / causal attention / SwiGLU run on T rows (prefill.cuh).  fp16 tensor-core
// operands: logits agree with the exact path to ~1e-3 relative, so the token-by-token entry points stay exact and this
// path is taken only by the batch entry point b200_prefill (B200_PREFILL_GEMM=0 turns it off).
// tokens per pass: large enough that the GEMM grids ((rows / 128) x (T / 256) CTAs) fill the 148 SMs even for the
// 1024-row k / v projections; 47 K floats of activations per token (Llama-3-8B) = 386 MB at 2048
static int prefill_chunk() { static int v = std::max(32, std::min(4096, env_int("B200_PREFILL_CHUNK", 2048))); return v; }


// Dequant-GEMM dispatch: the warp-specialised persistent kernel (gemm_umma2.cuh) when the matrix has a raw-tile tensor map
// (B200_GEMM2=0: the first kernel everywhere), else the first kernel (gemm_umma.cuh; TMA-fed only for passes of <= 64 rows).
static Umma2EncodeFn gemm_encode_fn() {
    static Umma2EncodeFn fn = []() -> Umma2EncodeFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (Umma2EncodeFn)f;
    }();
    return fn;
}
static bool gemm2_enabled() { static int v = env_int("B200_GEMM2", 1); return v != 0; }
static cudaError_t gemm_dispatch(UmmaParams& p, int n_sm, cudaStream_t st, uint64_t* launches = nullptr) {
    if (gemm2_enabled() && gemm_encode_fn() && umma2_eligible(p)) {
        // dequant groups: 2 for wide token tiles (a third changes nothing at T = 2048: 927 vs 944 TFLOP/s), 4 for the narrow tiles of
        // batched decode (batch 32: 2 groups 8.94, 3 groups 8.70 ms per step; with the graph and programmatic launches 3 groups 6.84,
        // 4 groups 6.61 -- 704 threads at 80 registers, 182 bytes of spills); B200_GEMM2_GROUPS forces one value (wide tiles: <= 3)
        static int ng_env = env_int("B200_GEMM2_GROUPS", 0);
        const int ng = ng_env ? std::max(2, std::min(umma2_tn(p.T) <= 64 ? 4 : 3, ng_env)) : (umma2_tn(p.T) <= 64 ? 4 : 2);
        if (launches) *launches += (p.k_split && !p.tile_cnt) ? 2 : 1;
        return umma2_launch(gemm_encode_fn(), p, n_sm, 227 * 1024 - 2048, st, ng);
    }
    if (p.T > env_int("B200_GEMM_TMA_MAX_T", 64)) p.tmap = nullptr;   // first kernel: large tiles keep the direct reads (measured, round 1)
    p.tile_cnt = nullptr;                                             // first kernel: separate reduce kernel
    if (launches) *launches += p.k_split ? 2 : 1;
    return umma_launch(p, st);
}

// TMA tensor maps for the dequant-GEMM's weight tiles: box = 128 rows x one 256-element block (stream_pitch(type, 1) bytes).
// Matrices whose rows are not 16-byte multiples keep the direct global reads.
static int umma_tmaps_build(b200_ctx* c) {
    if (c->pf_tmaps_built) return B200_OK;
    c->pf_tmaps_built = true;
    if (!env_int("B200_GEMM_TMA", 1)) return B200_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return B200_OK;
    }
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    std::vector<CUtensorMap> maps;
    auto add = [&](const DevTensor& w) {
        if (!w.present() || !umma_type_ok(w.type) || (w.row_bytes & 15) || ((uintptr_t)w.d & 15) || (w.ne[0] % 256)) return;
        const int pitch = stream_pitch(w.type, 1);
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)(w.row_bytes / 4), (cuuint64_t)w.ne[1]};
        const cuuint64_t strides[1] = {(cuuint64_t)w.row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)(pitch / 4), (cuuint32_t)kUmmaM};
        const cuuint32_t estr[2] = {1, 1};
        if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)w.d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
        c->pf_tmap_of[w.d] = (int)maps.size();
        maps.push_back(tm);
    };
    for (const Layer& L : c->layers) { add(L.wq); add(L.wk); add(L.wv); add(L.wo); add(L.gate); add(L.up); add(L.down); }
    add(c->output.present() ? c->output : c->token_embd);
    if (maps.empty()) return B200_OK;
    CU_ALLOC(cudaMalloc(&c->pf_tmaps, maps.size() * sizeof(CUtensorMap)));
    CU(cudaMemcpy(c->pf_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    return B200_OK;
}
static void umma_set_tmap(b200_ctx* c, UmmaParams& p) {
    p.tmap = nullptr;
    // (the first kernel only uses the map for passes of <= 64 rows: gemm_dispatch drops it for larger ones)
    auto it = c->pf_tmap_of.find(p.w);
    if (it == c->pf_tmap_of.end() || !c->pf_tmaps) return;
    p.tmap = (const CUtensorMap*)c->pf_tmaps + it->second;
    p.raw_pitch = stream_pitch(p.type, 1);
    p.raw_bytes = 256 / type_block_elems(p.type) * type_block_bytes(p.type);
}

// <<<grid, block, smem, st>>> with the programmatic-stream-serialization attribute when `pdl` (the kernel must call pdl_wait()
// before it touches anything: every kernel of the batched-decode pass does)
template <typename... KArgs, typename... Args>
static cudaError_t launch_on(cudaStream_t st, bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = pdl ? at : nullptr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// `prompt`: the caller is b200_prefill (consecutive positions of one sequence).  With the int8 KV cache only that case takes the GEMM
// path, and only with the tensor-core attention (its fp16 K / V^T tiles are built from the dequantised rows, kv_int8.cuh).
static bool prefill_gemm_ok(const b200_ctx* c, bool prompt = false) {
    const b200_model_desc& d = c->d;
    if (!c->use_prefill_gemm || c->par.world_size > 1 || d.n_experts > 0 || c->use_taps) return false;
    // ... and only on request (B200_KV_INT8_GEMM=1): the K / V rows of the GEMM pass differ from the exact path's by the fp16 operand
    // rounding (~1e-4) BEFORE they are quantised, the rounding to int8 codes is discontinuous (a neighbouring code is 1/127 of the
    // row maximum away), and the logits then sit 1e-3 .. 7e-3 from the oracle's (measured, tests/test_kv_int8.py) -- outside the
    // 1e-3 bound every default path keeps, so the default prompt path of an int8 context is the exact token-by-token one.
    if (c->kv_format == 1 && !(prompt && c->kv_int8_gemm && c->pf_attn_tc && gemm_encode_fn() && attn_umma_ok(d.head_dim, d.n_heads, d.n_kv_heads))) return false;
    if ((d.head_dim != 64 && d.head_dim != 128) || d.n_heads % d.n_kv_heads || d.n_heads / d.n_kv_heads > 8) return false;
    auto ok = [&](const DevTensor& w, int K) {
        UmmaParams p{};
        p.w = w.d; p.row_bytes = w.row_bytes; p.type = w.type; p.n_rows = (int)w.ne[1]; p.K = K; p.T = 1;
        p.x = reinterpret_cast<const __half*>(c->xa); p.ldx = K;   // pipeline buffers are 256-byte aligned
        return w.present() && (int)w.ne[0] == K && umma_eligible(p);
    };
    const int H = d.hidden, A = d.n_heads * d.head_dim, I = (int)d.ffn;
    if (H % 8 || A % 8 || I % 8) return false;
    for (const Layer& L : c->layers)
        if (!ok(L.wq, H) || !ok(L.wk, H) || !ok(L.wv, H) || !ok(L.wo, A) || !ok(L.gate, H) || !ok(L.up, H) || !ok(L.down, I)) return false;
    return true;
}

// seqs == nullptr: the n tokens are consecutive positions of slot `seq` (prefill).  seqs != nullptr: batched decode, token i
// is the next token of slot seqs[i] (distinct slots); the logits of all n rows are left in c->pf_logits ([n][vocab]).
static int prefill_gemm(b200_ctx* c, int seq, const uint32_t* tokens, int n, bool want_logits, const int* seqs = nullptr) {
    const b200_model_desc& d = c->d;
    Slot& sl = c->slots[seq];
    const int H = d.hidden, hd = d.head_dim, nh = d.n_heads, nkv = d.n_kv_heads, A = nh * hd, I = (int)d.ffn;
    const int QKV = (nh + 2 * nkv) * hd;
    const size_t per_tok = (size_t)H + QKV + 2 * (size_t)I + ((size_t)H + A + I + 1) / 2;   // floats (the fp16 rows count half)
    const int cap = std::min(prefill_chunk(), d.max_seq_len);
    if (!c->pf_buf) {
        CU_ALLOC(cudaMalloc((void**)&c->pf_buf, per_tok * cap * sizeof(float)));
        CU_ALLOC(cudaMalloc((void**)&c->pf_tok, cap * sizeof(int)));
        CU_ALLOC(cudaMalloc((void**)&c->pf_rows, (size_t)cap * 24));
    }
    { int rc_t = umma_tmaps_build(c); if (rc_t) return rc_t; }
    if (!c->pf_split) {   // split-K scratch for passes of <= 64 rows: 8 partial tiles of the widest projection
        c->pf_split_floats = (size_t)8 * 64 * std::max(std::max(I, QKV), H);
        CU_ALLOC(cudaMalloc((void**)&c->pf_split, c->pf_split_floats * sizeof(float)));
        CU_ALLOC(cudaMalloc((void**)&c->pf_tile_cnt, 8192 * sizeof(int)));
        CU(cudaMemset(c->pf_tile_cnt, 0, 8192 * sizeof(int)));
    }
    // batched decode: per-row (position, KV base of the slot, SeqState of the slot)
    const bool rows = seqs != nullptr;
    // tensor-core attention of a prompt chunk (attn_umma.cuh): fp16 K / V^T copies of ONE layer (rewritten per layer and chunk)
    bool attn_tc = !rows && c->pf_attn_tc && gemm_encode_fn() && attn_umma_ok(hd, nh, nkv);
    if (attn_tc && !c->pf_k16) {
        const int P = (d.max_seq_len + 127) & ~127;
        CU_ALLOC(cudaMalloc((void**)&c->pf_k16, (size_t)2 * nkv * P * hd * sizeof(__half)));    // hi rows, then lo rows
        CU_ALLOC(cudaMalloc((void**)&c->pf_vt16, (size_t)2 * nkv * P * hd * sizeof(__half)));   // hi rows, then lo rows
        c->pf_kv16_P = P;
        if (!attn_umma_encode(gemm_encode_fn(), c->pf_k16, c->pf_vt16, nkv, hd, P, &c->pf_kmap, &c->pf_vmap)) {
            c->pf_attn_tc = false;
            attn_tc = false;
        }
    }
    const bool q8 = c->kv_format == 1;
    if (q8 && (rows || !attn_tc)) return fail(B200_ERR_UNSUPPORTED, "int8 KV cache: the GEMM pass needs the tensor-core prompt attention");
    int* d_pos = reinterpret_cast<int*>(c->pf_rows);
    float** d_kv = reinterpret_cast<float**>(c->pf_rows + (size_t)cap * 8);
    SeqState** d_st = reinterpret_cast<SeqState**>(c->pf_rows + (size_t)cap * 16);
    if (rows) {
        if (n > cap) return fail(B200_ERR_INVALID_ARGUMENT, "batched decode: more rows than the activation buffers hold");
        std::vector<int> hp(n);
        std::vector<float*> hk(n);
        std::vector<SeqState*> hs(n);
        for (int i = 0; i < n; i++) {
            Slot& s = c->slots[seqs[i]];
            hp[i] = (int)s.host_pos; hk[i] = s.kv; hs[i] = s.d_state;
        }
        CU(cudaMemcpyAsync(d_pos, hp.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_kv, hk.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_st, hs.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));   // the host vectors go out of scope
        if (c->pf_logits_rows < n) {
            cudaFree(c->pf_logits);
            c->pf_logits = nullptr;
            CU_ALLOC(cudaMalloc((void**)&c->pf_logits, (size_t)n * d.vocab * sizeof(float)));
            c->pf_logits_rows = n;
            for (auto& bg : c->batch_graphs) {   // (the captured head GEMMs write the old buffer)
                if (bg.exec) cudaGraphExecDestroy(bg.exec);
                bg = b200_ctx::BatchGraph();
            }
        }
    }
    // CUDA graph of the step (rows mode): replay if this row count has one, capture on its second step (the first runs eagerly:
    // function attributes and lazy allocations happen there), and run the capture's graph to do the step itself
    struct CaptureGuard {
        cudaStream_t st;
        bool on = false;
        ~CaptureGuard() {
            if (!on) return;
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(st, &g);
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
        }
    } capg{c->stream};
    const bool use_graph = rows && c->batch_graph && !c->use_taps;
    const uint64_t launches0 = c->launches;
    if (use_graph) {
        if ((int)c->batch_graphs.size() <= n) c->batch_graphs.resize((size_t)n + 1);
        CU(cudaMemcpyAsync(c->pf_tok, tokens, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));   // the caller's buffer: not part of the graph
        b200_ctx::BatchGraph& bg = c->batch_graphs[n];
        if (bg.exec) {
            CU(cudaGraphLaunch(bg.exec, c->stream));
            c->launches += bg.launches;
            return B200_OK;
        }
        if (bg.warm) {
            CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            capg.on = true;
        } else {
            bg.warm = true;
        }
    }
    float* X = c->pf_buf;                               // [T][H] residual stream (f32)
    float* Q = X + (size_t)cap * H;                     // [T][QKV] f32
    float* G = Q + (size_t)cap * QKV;                   // [T][I] gate
    float* U = G + (size_t)cap * I;                     // [T][I] up
    __half* XNh = reinterpret_cast<__half*>(U + (size_t)cap * I);   // [T][H] normed rows, fp16 (GEMM input)
    __half* ATh = XNh + (size_t)cap * H;                // [T][A] attention output, fp16
    __half* Hh = ATh + (size_t)cap * A;                 // [T][I] silu(gate) * up, fp16
    const size_t kv_layer = (size_t)2 * nkv * d.max_seq_len * hd;
    cudaStream_t st = c->stream;
    const bool pdl = rows && c->batch_pdl;   // batched decode: programmatic dependent launches along the whole chain
    auto gemm = [&](const DevTensor& w, int K, const __half* x, int ldx, int T, float* y, int ldy, const DevTensor* bias, int acc) -> cudaError_t {
        UmmaParams p{};
        p.w = w.d; p.row_bytes = w.row_bytes; p.type = w.type; p.n_rows = (int)w.ne[1]; p.K = K;
        p.x = x; p.ldx = ldx; p.T = T; p.y = y; p.ldy = ldy;
        p.bias = (bias && bias->present()) ? bias->f32() : nullptr;
        p.accumulate = acc; p.err = c->mma_err; p.pdl = pdl ? 1 : 0;
        umma_set_tmap(c, p);
        if (gemm2_enabled() && gemm_encode_fn() && umma2_eligible(p) && env_int("B200_GEMM2_SPLIT_PLAN", 1)) umma2_plan_split(p, c->pf_split, c->pf_split_floats, c->n_sm);
        else umma_plan_split(p, c->pf_split, c->pf_split_floats, c->n_sm);
        if (p.k_split && (p.n_rows + kUmmaM - 1) / kUmmaM <= 8192 && env_int("B200_GEMM2_FUSED_REDUCE", 0)) p.tile_cnt = c->pf_tile_cnt;
        return gemm_dispatch(p, c->n_sm, st, &c->launches);
    };
    int last_T = 0;
    for (int done = 0; done < n; done += cap) {
        const int T = std::min(cap, n - done);
        const int pos0 = (int)sl.host_pos + done;
        last_T = T;
        if (!use_graph) CU(cudaMemcpyAsync(c->pf_tok, tokens + done, (size_t)T * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(launch_on(st, pdl, prefill_embed_kernel, dim3(T), dim3(256), 0, c->token_embd.type, (const uint8_t*)c->token_embd.d, c->token_embd.row_bytes, H, (const int*)c->pf_tok, d.vocab, X));
        for (int l = 0; l < d.n_layers; l++) {
            Layer& L = c->layers[l];
            float* kc = sl.kv + (size_t)l * kv_layer;
            float* vc = kc + kv_layer / 2;
            CU(launch_on(st, pdl, prefill_rms_norm_kernel, dim3(T), dim3(256), 0, (const float*)X, L.attn_norm.f32(), d.norm_eps, XNh, H));
            CU(gemm(L.wq, H, XNh, H, T, Q, QKV, &L.bq, 0));
            CU(gemm(L.wk, H, XNh, H, T, Q + A, QKV, &L.bk, 0));
            CU(gemm(L.wv, H, XNh, H, T, Q + A + nkv * hd, QKV, &L.bv, 0));
            PrefillRopeParams rp{};
            rp.qkv = Q; rp.ld = QKV; rp.k_cache = kc; rp.v_cache = vc; rp.freq = c->rope_freq; rp.pos0 = pos0;
            rp.n_heads = nh; rp.n_kv = nkv; rp.hd = hd; rp.max_seq = d.max_seq_len; rp.neox = d.rope_neox; rp.rope_scale = d.rope_scale;
            if (rows) { rp.row_pos = d_pos; rp.row_kv = d_kv; rp.k_off = (long long)((size_t)l * kv_layer); rp.v_off = rp.k_off + (long long)(kv_layer / 2); }
            const size_t q8_rows = (size_t)2 * nkv * d.max_seq_len;   // int8 KV: rows per layer (K then V), bytes = rows * hd
            signed char* k8 = q8 ? sl.kv8 + (size_t)l * q8_rows * hd : nullptr;
            signed char* v8 = q8 ? k8 + q8_rows * hd / 2 : nullptr;
            float* k8s = q8 ? sl.kv_scale + (size_t)l * q8_rows : nullptr;
            float* v8s = q8 ? k8s + q8_rows / 2 : nullptr;
            if (q8) {
                RopeKvQ8Params qp{};
                qp.q = Q; qp.k8 = k8; qp.v8 = v8; qp.k_scale = k8s; qp.v_scale = v8s; qp.freq = c->rope_freq;
                qp.n_heads = nh; qp.n_kv = nkv; qp.hd = hd; qp.max_seq = d.max_seq_len; qp.neox = d.rope_neox; qp.rope_scale = d.rope_scale;
                const dim3 rgrid((nh + 2 * nkv + 3) / 4, T);
                if (hd == 128) prefill_rope_kv_q8_kernel<4><<<rgrid, 128, 0, st>>>(qp, QKV, pos0);
                else prefill_rope_kv_q8_kernel<2><<<rgrid, 128, 0, st>>>(qp, QKV, pos0);
            } else {
                CU(launch_on(st, pdl, prefill_rope_kv_kernel, dim3(T, T <= 64 ? 4 : 1), dim3(256), 0, rp));   // few rows (batched decode): 4 CTAs per token
            }
            PrefillAttnParams ap{};
            ap.qkv = Q; ap.ld = QKV; ap.k_cache = kc; ap.v_cache = vc; ap.out = ATh; ap.ldo = A; ap.pos0 = pos0; ap.T = T;
            ap.n_heads = nh; ap.n_kv = nkv; ap.max_seq = d.max_seq_len; ap.scale = 1.0f / sqrtf((float)hd);
            if (rows) { ap.row_pos = d_pos; ap.row_kv = d_kv; ap.k_off = rp.k_off; ap.v_off = rp.v_off; }
            if (attn_tc) {
                const int kv_end = pos0 + T, kv_pad = (kv_end + 127) & ~127;
                AttnUmmaParams up{};
                up.qkv = Q; up.ld = QKV; up.out = ATh; up.ldo = A; up.pos0 = pos0; up.T = T; up.n_heads = nh; up.n_kv = nkv; up.G = nh / nkv;
                up.P = c->pf_kv16_P; up.scale = ap.scale; up.err = c->mma_err;
                const dim3 kgrid(kv_pad / 64, nkv);
                if (hd == 128) {
                    if (q8) prefill_kv16_q8_kernel<128><<<kgrid, 256, 0, st>>>(k8, v8, k8s, v8s, d.max_seq_len, kv_end, c->pf_kv16_P, c->pf_k16, c->pf_vt16);
                    else prefill_kv16_kernel<128><<<kgrid, 256, 0, st>>>(kc, vc, d.max_seq_len, kv_end, c->pf_kv16_P, c->pf_k16, c->pf_vt16);
                    CU(attn_umma_launch_hd<128>(c->pf_kmap, c->pf_vmap, up, st));
                } else {
                    if (q8) prefill_kv16_q8_kernel<64><<<kgrid, 256, 0, st>>>(k8, v8, k8s, v8s, d.max_seq_len, kv_end, c->pf_kv16_P, c->pf_k16, c->pf_vt16);
                    else prefill_kv16_kernel<64><<<kgrid, 256, 0, st>>>(kc, vc, d.max_seq_len, kv_end, c->pf_kv16_P, c->pf_k16, c->pf_vt16);
                    CU(attn_umma_launch_hd<64>(c->pf_kmap, c->pf_vmap, up, st));
                }
                c->launches += 1;
            } else {
                const int ablocks = (int)(((long long)T * nkv * 32 + 127) / 128);
                const int Gq = nh / nkv;
                if (hd == 128) {
                    CU(launch_on(st, pdl, Gq <= 4 ? prefill_attn_kernel<128, 4> : prefill_attn_kernel<128, 8>, dim3(ablocks), dim3(128), 0, ap));
                } else {
                    CU(launch_on(st, pdl, Gq <= 4 ? prefill_attn_kernel<64, 4> : prefill_attn_kernel<64, 8>, dim3(ablocks), dim3(128), 0, ap));
                }
            }
            CU(gemm(L.wo, A, ATh, A, T, X, H, nullptr, 1));                       // X += Wo attn
            CU(launch_on(st, pdl, prefill_rms_norm_kernel, dim3(T), dim3(256), 0, (const float*)X, L.ffn_norm.f32(), d.norm_eps, XNh, H));
            CU(gemm(L.gate, H, XNh, H, T, G, I, nullptr, 0));
            CU(gemm(L.up, H, XNh, H, T, U, I, nullptr, 0));
            CU(launch_on(st, pdl, prefill_swiglu_kernel, dim3(std::min(148 * 8, (int)(((long long)T * I + 255) / 256))), dim3(256), 0, (const float*)G, (const float*)U, Hh, (long long)T * I));
            CU(gemm(L.down, I, Hh, I, T, X, H, nullptr, 1));                      // X += Wd act
            c->launches += 5;
        }
        c->launches += 1;
        CU(cudaGetLastError());
    }
    if (rows) {   // every row's slot advances by one; final RMSNorm + vocab head of ALL rows through the GEMM
        CU(launch_on(st, pdl, prefill_advance_rows_kernel, dim3((n + 127) / 128), dim3(128), 0, (SeqState* const*)d_st, n));
        const DevTensor& head = c->output.present() ? c->output : c->token_embd;
        CU(launch_on(st, pdl, prefill_rms_norm_kernel, dim3(n), dim3(256), 0, (const float*)X, c->output_norm.f32(), d.norm_eps, XNh, H));
        UmmaParams hp{};
        hp.w = head.d; hp.row_bytes = head.row_bytes; hp.type = head.type; hp.n_rows = d.vocab; hp.K = H;
        hp.x = XNh; hp.ldx = H; hp.T = n; hp.y = c->pf_logits; hp.ldy = d.vocab; hp.err = c->mma_err; hp.pdl = pdl ? 1 : 0;
        umma_set_tmap(c, hp);
        CU(gemm_dispatch(hp, c->n_sm, st));
        c->launches += 3;
        if (capg.on) {
            capg.on = false;
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ex = nullptr;
            cudaError_t e = cudaStreamEndCapture(st, &g);
            if (e == cudaSuccess && g) e = cudaGraphInstantiate(&ex, g, 0);
            if (g) cudaGraphDestroy(g);
            if (e != cudaSuccess || !ex) {   // not capturable here: first without programmatic edges, then eager launches (nothing has run yet: do the step)
                cudaGetLastError();
                if (c->batch_pdl) c->batch_pdl = false;
                else c->batch_graph = false;
                c->launches = launches0;
                return prefill_gemm(c, seq, tokens, n, want_logits, seqs);
            }
            b200_ctx::BatchGraph& bg = c->batch_graphs[n];
            bg.exec = ex;
            bg.launches = c->launches - launches0;
            CU(cudaGraphLaunch(ex, st));
        }
        return B200_OK;
    }
    prefill_advance_kernel<<<1, 32, 0, st>>>(sl.d_state, n);
    if (want_logits) {   // final RMSNorm + vocab head of the last token, through the exact GEMV path
        CU(cudaMemcpyAsync(c->xa, X + (size_t)(last_T - 1) * H, (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, st));
        GemvParams p{};
        fill_seg(p.seg[0], c->output.present() ? c->output : c->token_embd, c->logits, nullptr, 0);
        p.n_seg = 1; p.K = H; p.x = c->xa; p.norm_w = c->output_norm.f32(); p.eps = d.norm_eps; p.epi = EPI_STORE;
        CU(launch_gemv(c, p));
    }
    c->prefill_gemm_tokens += (uint64_t)n;
    return B200_OK;
}

// Run one token: graph replay when enabled (captured lazily per slot and mode), else eager.
static int run_token(b200_ctx* c, int slot_i, Mode mode) {
    if (c->mega_ok)
        return mega_launch(c, slot_i, mode == MODE_LOGITS ? MEGA_LOGITS : mode == MODE_PREFILL ? MEGA_PREFILL : MEGA_GREEDY, 1);
    Slot& sl = c->slots[slot_i];
    const bool graph = c->use_graph && !c->use_taps;
    if (!graph) {
        cudaError_t e = enqueue_token(c, slot_i, mode);
        if (e != cudaSuccess) return fail(B200_ERR_OPERATION_FAILED, std::string("kernel launch: ") + cudaGetErrorString(e));
        return B200_OK;
    }
    if (!sl.graph[mode]) {
        uint64_t before = c->launches;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = enqueue_token(c, slot_i, mode);
        cudaGraph_t g = nullptr;
        cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
        sl.graph_launches[mode] = c->launches - before;
        c->launches = before;
        if (e != cudaSuccess || e2 != cudaSuccess || !g) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            if (c->use_pdl) {  // retry once without programmatic edges
                c->use_pdl = false;
                return run_token(c, slot_i, mode);
            }
            return fail(B200_ERR_OPERATION_FAILED, std::string("graph capture failed: ") +
                                                       cudaGetErrorString(e != cudaSuccess ? e : e2));
        }
        cudaError_t e3 = cudaGraphInstantiate(&sl.graph[mode], g, 0);
        cudaGraphDestroy(g);
        if (e3 != cudaSuccess) {
            sl.graph[mode] = nullptr;
            cudaGetLastError();
            if (c->use_pdl) {
                c->use_pdl = false;
                return run_token(c, slot_i, mode);
            }
            return fail(B200_ERR_OPERATION_FAILED, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e3));
        }
    }
    CU(cudaGraphLaunch(sl.graph[mode], c->stream));
    c->launches += sl.graph_launches[mode];
    return B200_OK;
}

// A speculative token that nobody will ask for: wait for it, then put the device position back to the last accepted token (the KV
// rows it wrote lie beyond the position and are overwritten by the next real token).
static int spec_drain(b200_ctx* c, int seq) {
    Slot& sl = c->slots[seq];
    if (!sl.spec) return B200_OK;
    sl.spec = false;
    c->spec_misses++;
    CU(cudaSetDevice(c->par.device));
    CU(cudaStreamSynchronize(c->stream));
    const int pos = (int)sl.host_pos;
    CU(cudaMemcpy(&sl.d_state->pos_next, &pos, sizeof(int), cudaMemcpyHostToDevice));
    return B200_OK;
}
static int check_slot(b200_ctx* c, int seq, const char* fn) {
    if (!c) return fail(B200_ERR_INVALID_ARGUMENT, std::string(fn) + ": null ctx");
    if (!c->finalized) return fail(B200_ERR_INVALID_ARGUMENT, std::string(fn) + ": context not finalized");
    if (seq < 0 || seq >= (int)c->slots.size()) return fail(B200_ERR_INVALID_ARGUMENT, std::string(fn) + ": bad sequence slot");
    return spec_drain(c, seq);   // (every entry point but b200_forward's hit path and b200_position starts from a settled slot)
}
static int check_token(b200_ctx* c, int seq, uint32_t token, const char* fn) {
    if (token >= (uint32_t)c->d.vocab)
        return fail(B200_ERR_INVALID_ARGUMENT, std::string(fn) + ": token id " + std::to_string(token) + " exceeds vocab size");
    if (c->slots[seq].host_pos + 1 > (uint64_t)c->d.max_seq_len)
        return fail(B200_ERR_INVALID_ARGUMENT, std::string(fn) + ": context length exceeded");
    return B200_OK;
}

static int set_token(b200_ctx* c, Slot& sl, uint32_t token) {
    *c->h_token = (int)token;
    CU(cudaMemcpyAsync(&sl.d_state->token, c->h_token, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    return B200_OK;
}

// The megakernels never hang: a wait that gives up (~1 s) raises the watchdog words and the launch finishes with garbage.
// Every synchronising entry point fetches the words with its results and turns them into an error (fail loudly).
#define WATCHDOG_FETCH(c) CU(cudaMemcpyAsync((c)->h_err, (c)->mma_err, 8 * sizeof(int), cudaMemcpyDeviceToHost, (c)->stream))
static int watchdog_check(b200_ctx* c, const char* who) {
    if (!c->h_err[0]) return B200_OK;
    char msg[200];
    snprintf(msg, sizeof msg, "%s: device watchdog tripped (code %d, wait %d, CTA %d, sequence %d): results discarded", who, c->h_err[0],
             c->h_err[1], c->h_err[2], c->h_err[3]);
    // reported once: clear the words (a later call starts clean) and resynchronise every slot's host position with the device's
    // (the launch may have advanced pos_next before it gave up)
    cudaMemset(c->mma_err, 0, 8 * sizeof(int));
    memset(c->h_err, 0, 8 * sizeof(int));
    for (Slot& sl : c->slots) {
        SeqState st{};
        if (cudaMemcpy(&st, sl.d_state, sizeof st, cudaMemcpyDeviceToHost) == cudaSuccess) sl.host_pos = (uint64_t)std::max(st.pos_next, 0);
    }
    return fail(B200_ERR_OPERATION_FAILED, msg);
}

// GpuModelWrapper::forward prefills a prompt by calling prefill_token once per token (src/backend/mod.rs:343-346), which would keep
// the tcgen05 GEMM prefill out of reach of an unchanged engine.  With B200_PREFILL_QUEUE=1 b200_prefill_token only queues the token
// (position() counts it); the queue runs as one b200_prefill at the next call that needs the sequence state.  Off by default: the
// GEMM path rounds its operands to fp16 (logits within 3e-3 of the exact path), the token-by-token path is exact.
static int flush_pending(b200_ctx* c, int seq) {
    Slot& sl = c->slots[seq];
    if (sl.pending.empty()) return B200_OK;
    std::vector<uint32_t> toks;
    toks.swap(sl.pending);
    c->queue_bypass = true;
    const int rc = b200_prefill(c, seq, toks.data(), (int)toks.size(), nullptr);
    c->queue_bypass = false;
    return rc;
}

extern "C" int b200_ctx_set_kv_format(b200_ctx* c, int format) {
    if (!c) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_set_kv_format: null ctx");
    if (c->finalized) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_set_kv_format: the context is finalized already");
    if (format != 0 && format != 1) return fail(B200_ERR_UNSUPPORTED, "b200_ctx_set_kv_format: 0 = f32, 1 = int8 (FP8 formats of kv_quantized.rs are not built)");
    c->kv_format = format;
    return B200_OK;
}
extern "C" int b200_ctx_kv_format(b200_ctx* c, int* out) {
    if (!c || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_kv_format: bad argument");
    *out = c->kv_format;
    return B200_OK;
}

// argmax of the logits just produced -> st->token, its value to the host, the next token launched on it, its logits to h_spec_logits[buf]
static int spec_enqueue(b200_ctx* c, int seq, int buf) {
    Slot& sl = c->slots[seq];
    if (!c->speculate || !c->mega_ok || c->par.world_size > 1 || c->use_taps || sl.host_pos + 2 > (uint64_t)c->d.max_seq_len) return B200_OK;
    argmax_kernel<<<1, 1024, 0, c->stream>>>((const float*)c->logits, c->d.vocab, sl.d_state, sl.d_generated, 0);
    CU(cudaMemcpyAsync(&c->h_spec_pick[buf], &sl.d_state->token, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    int rc;
    if ((rc = mega_launch(c, seq, MEGA_LOGITS, 1))) return rc;
    CU(cudaMemcpyAsync(c->h_spec_logits[buf], c->logits, (size_t)c->vocab_l * 4, cudaMemcpyDeviceToHost, c->stream));
    WATCHDOG_FETCH(c);
    CU(cudaEventRecord(c->spec_ev[buf], c->stream));
    c->launches++;
    sl.spec = true;
    sl.spec_buf = buf;
    return B200_OK;
}

extern "C" int b200_forward(b200_ctx* c, int seq, uint32_t token, float* logits_out) {
    int rc;
    if (c && c->finalized && seq >= 0 && seq < (int)c->slots.size() && c->slots[seq].spec && logits_out) {
        Slot& sl = c->slots[seq];
        const int buf = sl.spec_buf;
        CU(cudaSetDevice(c->par.device));
        CU(cudaEventSynchronize(c->spec_ev[buf]));
        if ((uint32_t)c->h_spec_pick[buf] == token && !c->h_err[0]) {   // hit: the token in flight IS this call's token
            sl.spec = false;
            sl.host_pos++;
            c->spec_hits++;
            if ((rc = spec_enqueue(c, seq, buf ^ 1))) return rc;        // keep the device busy while the host copies and picks
            memcpy(logits_out, c->h_spec_logits[buf], (size_t)c->vocab_l * 4);
            return B200_OK;
        }
    }
    if ((rc = check_slot(c, seq, "b200_forward"))) return rc;               // (drains a speculation that missed)
    if (!logits_out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_forward: null logits_out");
    if ((rc = flush_pending(c, seq))) return rc;
    if ((rc = check_token(c, seq, token, "b200_forward"))) return rc;
    CU(cudaSetDevice(c->par.device));
    Slot& sl = c->slots[seq];
    if ((rc = set_token(c, sl, token))) return rc;
    if ((rc = run_token(c, seq, MODE_LOGITS))) return rc;
    if (c->speculate && c->mega_ok && c->par.world_size == 1 && !c->use_taps) {
        for (int o = 0; o < (int)c->slots.size(); o++)      // the staging buffers are the context's: one sequence runs ahead at a time
            if (o != seq && (rc = spec_drain(c, o))) return rc;
        CU(cudaMemcpyAsync(c->h_spec_logits[0], c->logits, (size_t)c->vocab_l * 4, cudaMemcpyDeviceToHost, c->stream));
        WATCHDOG_FETCH(c);
        CU(cudaEventRecord(c->spec_ev[0], c->stream));
        sl.host_pos++;                                                       // (spec_enqueue looks at the position after this token)
        if ((rc = spec_enqueue(c, seq, 1))) { sl.host_pos--; return rc; }
        CU(cudaEventSynchronize(c->spec_ev[0]));
        if ((rc = watchdog_check(c, "b200_forward"))) { sl.host_pos--; return rc; }
        memcpy(logits_out, c->h_spec_logits[0], (size_t)c->vocab_l * 4);
        return B200_OK;
    }
    // tensor parallel: this rank's slice [rank * vocab/P, (rank+1) * vocab/P) of the logits (the caller gathers)
    CU(cudaMemcpyAsync(c->h_logits, c->logits, (size_t)c->vocab_l * 4, cudaMemcpyDeviceToHost, c->stream));
    WATCHDOG_FETCH(c);
    CU(cudaStreamSynchronize(c->stream));
    if ((rc = watchdog_check(c, "b200_forward"))) return rc;
    memcpy(logits_out, c->h_logits, (size_t)c->vocab_l * 4);
    sl.host_pos++;
    return B200_OK;
}

extern "C" int b200_prefill_token(b200_ctx* c, int seq, uint32_t token) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_prefill_token"))) return rc;
    if ((rc = check_token(c, seq, token, "b200_prefill_token"))) return rc;
    CU(cudaSetDevice(c->par.device));
    Slot& sl = c->slots[seq];
    if (c->prefill_queue && !c->queue_bypass && prefill_gemm_ok(c)) {
        if (sl.host_pos + sl.pending.size() + 1 > (uint64_t)c->d.max_seq_len)
            return fail(B200_ERR_INVALID_ARGUMENT, "b200_prefill_token: context length exceeded");
        sl.pending.push_back(token);
        return B200_OK;
    }
    if ((rc = set_token(c, sl, token))) return rc;
    if ((rc = run_token(c, seq, MODE_PREFILL))) return rc;
    WATCHDOG_FETCH(c);
    CU(cudaStreamSynchronize(c->stream));  // h_token is reused by the next call
    if ((rc = watchdog_check(c, "b200_prefill_token"))) return rc;
    sl.host_pos++;
    return B200_OK;
}

extern "C" int b200_prefill(b200_ctx* c, int seq, const uint32_t* tokens, int n, float* logits_out) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_prefill"))) return rc;
    if (!tokens || n <= 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_prefill: no tokens");
    if (!c->slots[seq].pending.empty() && (rc = flush_pending(c, seq))) return rc;
    if (n >= c->prefill_gemm_min && prefill_gemm_ok(c, true)) {
        for (int i = 0; i < n; i++)
            if (tokens[i] >= (uint32_t)c->d.vocab) return fail(B200_ERR_INVALID_ARGUMENT, "b200_prefill: token id exceeds vocab size");
        Slot& sl = c->slots[seq];
        if (sl.host_pos + (uint64_t)n > (uint64_t)c->d.max_seq_len) return fail(B200_ERR_INVALID_ARGUMENT, "b200_prefill: context length exceeded");
        CU(cudaSetDevice(c->par.device));
        if ((rc = prefill_gemm(c, seq, tokens, n, logits_out != nullptr))) return rc;
        if (logits_out) CU(cudaMemcpyAsync(c->h_logits, c->logits, (size_t)c->vocab_l * 4, cudaMemcpyDeviceToHost, c->stream));
        WATCHDOG_FETCH(c);
        CU(cudaStreamSynchronize(c->stream));
        if ((rc = watchdog_check(c, "b200_prefill"))) return rc;
        if (logits_out) memcpy(logits_out, c->h_logits, (size_t)c->vocab_l * 4);
        sl.host_pos += (uint64_t)n;
        return B200_OK;
    }
    for (int i = 0; i < n; i++) {
        if (i == n - 1 && logits_out) return b200_forward(c, seq, tokens[i], logits_out);
        if ((rc = b200_prefill_token(c, seq, tokens[i]))) return rc;
    }
    return B200_OK;
}

extern "C" int b200_decode_batch(b200_ctx* c, const int* seqs, const uint32_t* tokens, int n, float* logits_out) {
    if (!c || !seqs || !tokens || !logits_out || n <= 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_batch: bad argument");
    if (c->par.world_size > 1)   // a rank produces vocab / world_size logits per row: the n x vocab layout of this call does not apply
        return fail(B200_ERR_UNSUPPORTED, "b200_decode_batch: not available under tensor parallelism (use b200_forward per sequence)");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (seqs[i] == seqs[j]) return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_batch: duplicate sequence slot");
    int rc;
    for (int i = 0; i < n; i++)
        if (seqs[i] >= 0 && seqs[i] < (int)c->slots.size() && (rc = flush_pending(c, seqs[i]))) return rc;
    // n >= batch_gemm_min rows of an eligible dense model: ONE pass of the tcgen05 dequant-GEMMs for all the sequences (the
    // weights are read once per step instead of once per sequence); fp16 tensor-core operands, see gemm_umma.cuh
    if (n >= c->batch_gemm_min && prefill_gemm_ok(c)) {
        const DevTensor& head = c->output.present() ? c->output : c->token_embd;
        UmmaParams hp{};
        hp.w = head.d; hp.row_bytes = head.row_bytes; hp.type = head.type; hp.n_rows = c->d.vocab; hp.K = c->d.hidden; hp.T = 1;
        hp.x = reinterpret_cast<const __half*>(c->xa); hp.ldx = c->d.hidden;
        if (umma_eligible(hp) && n <= std::min(prefill_chunk(), c->d.max_seq_len)) {
            for (int i = 0; i < n; i++) {
                if ((rc = check_slot(c, seqs[i], "b200_decode_batch"))) return rc;
                if ((rc = check_token(c, seqs[i], tokens[i], "b200_decode_batch"))) return rc;
            }
            CU(cudaSetDevice(c->par.device));
            if ((rc = prefill_gemm(c, seqs[0], tokens, n, false, seqs))) return rc;
            WATCHDOG_FETCH(c);
            CU(cudaStreamSynchronize(c->stream));
            if ((rc = watchdog_check(c, "b200_decode_batch"))) return rc;
            {   // n x vocab logits (16.4 MB at batch 32 on Llama-3) to the caller's pageable buffer: DMA into a pinned buffer, rows copied
                // out by a few host threads (a pageable cudaMemcpy stages the same bytes through the driver on one thread)
                const size_t bytes = (size_t)n * c->d.vocab * sizeof(float);
                if (c->h_batch_logits_bytes < bytes) {
                    if (c->h_batch_logits) cudaFreeHost(c->h_batch_logits);
                    c->h_batch_logits = nullptr;
                    c->h_batch_logits_bytes = 0;
                    if (cudaHostAlloc((void**)&c->h_batch_logits, bytes, cudaHostAllocDefault) == cudaSuccess) c->h_batch_logits_bytes = bytes;
                    else cudaGetLastError();
                }
                if (c->h_batch_logits) {
                    CU(cudaMemcpy(c->h_batch_logits, c->pf_logits, bytes, cudaMemcpyDeviceToHost));
                    gguf_parallel_copy((uint8_t*)logits_out, (const uint8_t*)c->h_batch_logits, bytes, 4);
                } else {
                    CU(cudaMemcpy(logits_out, c->pf_logits, bytes, cudaMemcpyDeviceToHost));
                }
            }
            for (int i = 0; i < n; i++) c->slots[seqs[i]].host_pos++;
            return B200_OK;
        }
    }
    for (int i = 0; i < n; i++)
        if ((rc = b200_forward(c, seqs[i], tokens[i], logits_out + (size_t)i * c->d.vocab))) return rc;
    return B200_OK;
}

extern "C" int b200_reset(b200_ctx* c, int seq) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_reset"))) return rc;
    CU(cudaSetDevice(c->par.device));
    Slot& sl = c->slots[seq];
    sl.pending.clear();
    CU(cudaMemsetAsync(sl.d_state, 0, sizeof(SeqState), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    sl.host_pos = 0;
    return B200_OK;
}

extern "C" int b200_position(b200_ctx* c, int seq, uint64_t* out) {
    if (!c || !c->finalized || seq < 0 || seq >= (int)c->slots.size()) return fail(B200_ERR_INVALID_ARGUMENT, "b200_position: bad argument");
    if (!out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_position: null out");   // (host-side count of accepted tokens: a token in flight does not count)
    *out = c->slots[seq].host_pos + c->slots[seq].pending.size();   // queued prefill tokens count (GpuInference::position)
    return B200_OK;
}

extern "C" int b200_decode_greedy(b200_ctx* c, int seq, uint32_t first_token, int n_steps, uint32_t* tokens_out,
                                  float* elapsed_ms) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_decode_greedy"))) return rc;
    if (n_steps <= 0 || n_steps > kMaxGenerated) return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_greedy: bad n_steps");
    if (first_token >= (uint32_t)c->d.vocab) return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_greedy: token exceeds vocab size");
    if ((rc = flush_pending(c, seq))) return rc;
    Slot& sl = c->slots[seq];
    if (sl.host_pos + (uint64_t)n_steps > (uint64_t)c->d.max_seq_len)
        return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_greedy: context length exceeded");
    CU(cudaSetDevice(c->par.device));
    if ((rc = set_token(c, sl, first_token))) return rc;
    CU(cudaMemsetAsync(&sl.d_state->n_generated, 0, sizeof(int), c->stream));
    struct Events {   // destroyed on every exit path
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    } ev;
    CU(cudaEventCreate(&ev.e0));
    CU(cudaEventCreate(&ev.e1));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventRecord(ev.e0, c->stream));
    if (c->mega_ok) {
        if ((rc = mega_launch(c, seq, MEGA_GREEDY, n_steps))) return rc;
    } else {
        for (int i = 0; i < n_steps; i++)
            if ((rc = run_token(c, seq, MODE_GREEDY))) return rc;
    }
    CU(cudaEventRecord(ev.e1, c->stream));
    WATCHDOG_FETCH(c);
    CU(cudaStreamSynchronize(c->stream));
    if ((rc = watchdog_check(c, "b200_decode_greedy"))) return rc;
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    if (elapsed_ms) *elapsed_ms = ms;
    sl.host_pos += (uint64_t)n_steps;
    if (tokens_out) CU(cudaMemcpy(tokens_out, sl.d_generated, (size_t)n_steps * sizeof(int), cudaMemcpyDeviceToHost));
    return B200_OK;
}

// ------------------------------------------------------------------ single-process group: one context and one host thread per device
// The per-token kernels of the ranks wait for each other on the device, so the ranks' calls must be in flight together: every group call
// is handed to the persistent worker thread of each rank and returns when all of them have finished.
struct GroupWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false;
    int rc = 0;
    std::string msg;
};
struct b200_group {
    std::vector<b200_ctx*> ctx;
    std::vector<std::unique_ptr<GroupWorker>> workers;
    bool ep = false;
    int vocab = 0, vocab_l = 0;
};
static void group_worker_loop(GroupWorker* w) {
    for (;;) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job = w->job;
            w->has_job = false;
        }
        const int rc = job();
        const char* m = rc ? b200_last_error() : "";
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->rc = rc;
            w->msg = m ? m : "";
            w->done = true;
        }
        w->cv.notify_all();
    }
}
// runs fn(rank) on every rank's thread; the first failure (lowest rank) is returned with its message
template <class F>
static int group_run(b200_group* g, F fn) {
    const int n = (int)g->ctx.size();
    for (int r = 0; r < n; r++) {
        GroupWorker* w = g->workers[r].get();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->job = [fn, r]() { return fn(r); };
            w->has_job = true;
            w->done = false;
        }
        w->cv.notify_all();
    }
    int rc = B200_OK;
    std::string msg;
    for (int r = 0; r < n; r++) {
        GroupWorker* w = g->workers[r].get();
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc && !rc) { rc = w->rc; msg = w->msg; }
    }
    return rc ? fail(rc, msg) : B200_OK;
}

extern "C" void b200_group_destroy(b200_group* g) {
    if (!g) return;
    for (auto& w : g->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    for (b200_ctx* c : g->ctx) {
        if (!c) continue;
        for (int r = 0; r < c->par.world_size && r < kMmaMaxPeers; r++) c->tp_peer[r] = (r == c->par.rank) ? c->tp_peer[r] : nullptr;   // (not IPC mappings)
        b200_ctx_destroy(c);
    }
    delete g;
}

extern "C" int b200_group_create(const b200_model_desc* desc, int n_devices, const int* devices, b200_group** out) {
    if (!desc || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_create: null argument");
    if (n_devices != 1 && n_devices != 2 && n_devices != 4 && n_devices != 8) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_create: n_devices must be 1, 2, 4 or 8");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(B200_ERR_NOT_AVAILABLE, "cuda-b200: no CUDA device (this backend has no CPU fallback)"); }
    std::vector<int> dev(n_devices);
    for (int r = 0; r < n_devices; r++) {
        dev[r] = devices ? devices[r] : r;
        if (dev[r] < 0 || dev[r] >= n_dev) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_create: device ordinal out of range");
        for (int q = 0; q < r; q++) if (dev[q] == dev[r]) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_create: duplicate device");
    }
    b200_group* g = new b200_group();
    g->ctx.assign(n_devices, nullptr);
    for (int r = 0; r < n_devices; r++) {
        b200_parallel_desc par{n_devices, r, dev[r]};
        const int rc = b200_ctx_create(desc, &par, &g->ctx[r]);
        if (rc) { b200_group_destroy(g); return rc; }
    }
    g->ep = g->ctx[0]->ep;
    g->vocab = g->ctx[0]->d.vocab;
    g->vocab_l = g->ctx[0]->vocab_l;
    if (n_devices > 1) {
        // peer access both ways, then every rank's exchange region as every other rank sees it: the plain device pointer
        for (int r = 0; r < n_devices; r++) {
            cudaSetDevice(dev[r]);
            for (int q = 0; q < n_devices; q++) {
                if (q == r) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, dev[r], dev[q]);
                if (!can) { b200_group_destroy(g); return fail(B200_ERR_UNSUPPORTED, "b200_group_create: no peer access between the devices"); }
                const cudaError_t e = cudaDeviceEnablePeerAccess(dev[q], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { b200_group_destroy(g); return fail(B200_ERR_INITIALIZATION_FAILED, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
                cudaGetLastError();
            }
        }
        uint8_t h[64];
        for (int r = 0; r < n_devices; r++) {   // (allocates the region)
            const int rc = b200_ctx_tp_handle(g->ctx[r], h);
            if (rc) { b200_group_destroy(g); return rc; }
        }
        for (int r = 0; r < n_devices; r++)
            for (int q = 0; q < n_devices; q++) {
                g->ctx[r]->tp_peer[q] = g->ctx[q]->tp_region;
                g->ctx[r]->tp_peer_set[q] = true;
            }
    }
    for (int r = 0; r < n_devices; r++) {
        g->workers.emplace_back(new GroupWorker());
        GroupWorker* w = g->workers.back().get();
        w->th = std::thread(group_worker_loop, w);
    }
    *out = g;
    return B200_OK;
}

extern "C" int b200_group_upload_tensor(b200_group* g, const char* gguf_name, uint32_t ggml_type, const uint64_t* ne, int n_dims,
                                        const void* host, size_t nbytes) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_upload_tensor: null group");
    return group_run(g, [=](int r) { return b200_ctx_upload_tensor(g->ctx[r], gguf_name, ggml_type, ne, n_dims, host, nbytes); });
}
extern "C" int b200_group_finalize(b200_group* g) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_finalize: null group");
    return group_run(g, [=](int r) { return b200_ctx_finalize(g->ctx[r]); });
}
extern "C" int b200_group_forward(b200_group* g, int seq, uint32_t token, float* logits_out) {
    if (!g || !logits_out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_forward: null argument");
    if (g->ctx.size() == 1) return b200_forward(g->ctx[0], seq, token, logits_out);
    // tensor parallel: rank r produces rows [r * vocab / P, (r + 1) * vocab / P) of the logits, straight into the caller's row;
    // expert parallel: the head is replicated -- rank 0's row is the answer, the others land in their contexts' scratch rows
    return group_run(g, [=](int r) {
        if (g->ep && r > 0) return b200_prefill_token(g->ctx[r], seq, token);   // (same token, same layers; its logits are not needed)
        return b200_forward(g->ctx[r], seq, token, logits_out + (g->ep ? 0 : (size_t)r * g->vocab_l));
    });
}
extern "C" int b200_group_prefill_token(b200_group* g, int seq, uint32_t token) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_prefill_token: null group");
    return group_run(g, [=](int r) { return b200_prefill_token(g->ctx[r], seq, token); });
}
extern "C" int b200_group_reset(b200_group* g, int seq) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_reset: null group");
    return group_run(g, [=](int r) { return b200_reset(g->ctx[r], seq); });
}
extern "C" int b200_group_position(b200_group* g, int seq, uint64_t* out) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_position: null group");
    return b200_position(g->ctx[0], seq, out);
}
extern "C" int b200_group_decode_greedy(b200_group* g, int seq, uint32_t first_token, int n_steps, uint32_t* tokens_out, float* elapsed_ms) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_decode_greedy: null group");
    std::vector<float> ms(g->ctx.size(), 0.0f);
    float* msp = ms.data();
    const int rc = group_run(g, [=](int r) { return b200_decode_greedy(g->ctx[r], seq, first_token, n_steps, r == 0 ? tokens_out : nullptr, msp + r); });
    if (elapsed_ms) *elapsed_ms = *std::max_element(ms.begin(), ms.end());   // device time, max over the ranks
    return rc;
}
extern "C" int b200_group_size(b200_group* g, int* out) {
    if (!g || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_size: null argument");
    *out = (int)g->ctx.size();
    return B200_OK;
}
extern "C" int b200_group_ctx(b200_group* g, int rank, b200_ctx** out) {
    if (!g || !out || rank < 0 || rank >= (int)g->ctx.size()) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_ctx: bad argument");
    *out = g->ctx[rank];
    return B200_OK;
}

extern "C" int b200_get_hidden(b200_ctx* c, int seq, int layer, float* out) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_get_hidden"))) return rc;
    if (!c->use_taps) return fail(B200_ERR_UNSUPPORTED, "b200_get_hidden: taps are off (set B200_TAPS=1 before b200_ctx_create)");
    if (layer < 0 || layer > c->d.n_layers || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_get_hidden: bad layer");
    CU(cudaSetDevice(c->par.device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out, c->taps + (size_t)layer * c->d.hidden, (size_t)c->d.hidden * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// Debug: globaltimer stamps (ns) of CTA 0 after the embedding barrier and after every phase barrier of the last
// token run by the megakernel.  First call arms the stamps; returns the number of values written to out.
extern "C" int b200_debug_mega_timeline(b200_ctx* c, unsigned long long* out, int max_n) {
    if (!c || !c->mega_ok) return 0;
    cudaSetDevice(c->par.device);
    const int n = (c->stream2_ok ? c->s2_phases : c->mega_phases) + 1;
    if (!c->mega_dbg) {
        if (cudaMalloc((void**)&c->mega_dbg, (size_t)(n + 2) * 8) != cudaSuccess) return 0;
        cudaMemset(c->mega_dbg, 0, (size_t)(n + 2) * 8);
        return 0;
    }
    cudaStreamSynchronize(c->stream);
    const int m = std::min(n, max_n);
    if (out) cudaMemcpy(out, c->mega_dbg, (size_t)m * 8, cudaMemcpyDeviceToHost);
    return m;
}

// Debug: per-warp globaltimer stamps (gemv_mma.cuh MMA_STAMP) of one GEMV phase of slot 0's megakernel program.
// phase >= 0 arms (patches the phase descriptor); phase < 0 fetches grid*16*8 values into out.
extern "C" int b200_debug_mega_phase(b200_ctx* c, int phase, unsigned long long* out, int max_n) {
    if (!c || !c->mega_ok) return 0;
    cudaSetDevice(c->par.device);
    static unsigned long long* buf = nullptr;
    const size_t n = (size_t)c->n_sm * 16 * 8;
    if (!buf && cudaMalloc((void**)&buf, n * 8) != cudaSuccess) return 0;
    cudaStreamSynchronize(c->stream);
    if (phase >= 0) {
        MegaPhase* prog = c->stream2_ok ? c->slots[0].d_phases2 : c->slots[0].d_phases;
        if (phase >= (c->stream2_ok ? c->s2_phases : c->mega_phases)) return 0;
        cudaMemset(buf, 0, n * 8);
        MegaPhase ph;
        cudaMemcpy(&ph, prog + phase, sizeof ph, cudaMemcpyDeviceToHost);
        if (ph.kind == PH_GEMV) ph.gemv.dbg = buf;
        else if (ph.kind == PH_ATTN) ph.attn.dbg = buf;
        cudaMemcpy(prog + phase, &ph, sizeof ph, cudaMemcpyHostToDevice);
        return 1;
    }
    const int m = (int)std::min<size_t>(n, (size_t)max_n);
    if (out) cudaMemcpy(out, buf, (size_t)m * 8, cudaMemcpyDeviceToHost);
    return m;
}

// Debug: copy one of the activation buffers to the host (0 xa, 1 xb, 2 qkv, 3 attn, 4 hbuf, 5 logits).
extern "C" int b200_debug_read(b200_ctx* c, int which, float* out, int n) {
    if (!c || !out) return 0;
    cudaSetDevice(c->par.device);
    cudaStreamSynchronize(c->stream);
    const float* src[6] = {c->xa, c->xb, c->qkv, c->attn, c->hbuf, c->logits};
    if (which < 0 || which > 5) return 0;
    return cudaMemcpy(out, src[which], (size_t)n * 4, cudaMemcpyDeviceToHost) == cudaSuccess ? n : 0;
}

extern "C" int b200_ctx_stats(b200_ctx* c, uint64_t* kernel_launches, uint64_t* weight_bytes, uint64_t* kv_bytes_per_pos) {
    if (!c) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_stats: null ctx");
    if (kernel_launches) *kernel_launches = c->launches;
    if (weight_bytes) *weight_bytes = c->weight_bytes_per_token;
    if (kv_bytes_per_pos)   // f32 rows, or int8 rows + one f32 scale each
        *kv_bytes_per_pos = (uint64_t)2 * c->d.n_layers * c->d.n_kv_heads * (c->kv_format == 1 ? c->d.head_dim + 4 : c->d.head_dim * 4);
    return B200_OK;
}

// Which decode path finalize selected: 0 = CUDA graph of per-op kernels, 1 = per-token megakernel (mega.cuh),
// 2 = streamed megakernel (stream.cuh), 3 = second streamed megakernel (stream2.cuh).
extern "C" int b200_ctx_path(b200_ctx* c, int* out) {
    if (!c || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_path: null argument");
    *out = c->mega_ok ? (c->stream2_ok ? 3 : c->stream_ok ? 2 : 1) : 0;
    return B200_OK;
}

// Debug: the watchdog words of the tensor-pipe / megakernel paths (0 = no wait ever gave up); clears them.
extern "C" int b200_debug_err(b200_ctx* c, int* out8) {
    if (!c || !out8) return fail(B200_ERR_INVALID_ARGUMENT, "b200_debug_err: null argument");
    CU(cudaSetDevice(c->par.device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out8, c->mma_err, 8 * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemset(c->mma_err, 0, 8 * sizeof(int)));
    memset(c->h_err, 0, 8 * sizeof(int));
    return B200_OK;
}

// Roofline probe for bench.py: replays ONLY the gemv_kernel launches of one token (same
// arguments, order and launch attributes as the decode graph; embedding / RoPE / attention /
// argmax launches left out) `iters` times between two CUDA events on the launching stream.
extern "C" int b200_bench_gemv_pass(b200_ctx* c, int seq, int iters, float* avg_ms_per_pass, uint64_t* launches_per_pass,
                                    uint64_t* bytes_per_pass) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_bench_gemv_pass"))) return rc;
    if (iters <= 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_bench_gemv_pass: iters <= 0");
    CU(cudaSetDevice(c->par.device));
    if (!c->gemv_graph) {
        uint64_t before = c->launches;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = enqueue_token(c, seq, MODE_LOGITS, true);
        cudaGraph_t g = nullptr;
        cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
        c->gemv_graph_launches = c->launches - before;
        c->launches = before;
        if (e != cudaSuccess || e2 != cudaSuccess || !g) {
            cudaGetLastError();
            return fail(B200_ERR_OPERATION_FAILED, "b200_bench_gemv_pass: capture failed");
        }
        cudaError_t e3 = cudaGraphInstantiate(&c->gemv_graph, g, 0);
        cudaGraphDestroy(g);
        if (e3 != cudaSuccess) return fail(B200_ERR_OPERATION_FAILED, "b200_bench_gemv_pass: instantiate failed");
    }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaGraphLaunch(c->gemv_graph, c->stream));  // warm-up
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventRecord(e0, c->stream));
    for (int i = 0; i < iters; i++) CU(cudaGraphLaunch(c->gemv_graph, c->stream));
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->launches += c->gemv_graph_launches * (uint64_t)(iters + 1);
    if (avg_ms_per_pass) *avg_ms_per_pass = ms / iters;
    if (launches_per_pass) *launches_per_pass = c->gemv_graph_launches;
    if (bytes_per_pass) *bytes_per_pass = c->weight_bytes_per_token - c->token_embd.row_bytes;
    return B200_OK;
}

// Per-kernel GB/s: times plain vec_mat_q launches over resident weights.  A name with
// "%d" walks every layer's tensor in order (footprint >> L2, the decode access pattern);
// a plain name times one tensor and writes a 256 MB buffer between launches to flush L2.
extern "C" int b200_bench_weight_gemv(b200_ctx* c, const char* gguf_name, int iters, float* avg_ms, uint64_t* bytes) {
    if (!c || !c->finalized || !gguf_name || iters <= 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_bench_weight_gemv: bad argument");
    CU(cudaSetDevice(c->par.device));
    std::vector<const DevTensor*> ts;
    if (strstr(gguf_name, "%d")) {
        for (int l = 0; l < c->d.n_layers; l++) {
            char nm[128];
            snprintf(nm, sizeof nm, gguf_name, l);
            auto it = c->tensors.find(nm);
            if (it == c->tensors.end()) return fail(B200_ERR_INVALID_ARGUMENT, std::string("no tensor ") + nm);
            ts.push_back(&it->second);
        }
    } else {
        auto it = c->tensors.find(gguf_name);
        if (it == c->tensors.end()) return fail(B200_ERR_INVALID_ARGUMENT, std::string("no tensor ") + gguf_name);
        ts.push_back(&it->second);
    }
    uint64_t total = 0;
    size_t max_k = 0;
    for (const DevTensor* t : ts) {
        if (t->n_dims != 2) return fail(B200_ERR_UNSUPPORTED, "b200_bench_weight_gemv: 2-D weights only");
        if (t->ne[1] > c->out_scratch_elems) return fail(B200_ERR_UNSUPPORTED, "b200_bench_weight_gemv: output too large");
        total += t->nbytes;
        max_k = std::max<size_t>(max_k, t->ne[0]);
    }
    float* xin = nullptr;
    CU_ALLOC(cudaMalloc((void**)&xin, max_k * 4));
    CU(cudaMemset(xin, 0, max_k * 4));
    const bool flush = ts.size() == 1;
    if (flush && !c->flush_buf) {
        c->flush_bytes = 256u << 20;
        CU_ALLOC(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const bool pdl = c->use_pdl;
    c->use_pdl = false;
    double sum_ms = 0.0;
    for (int it = -1; it < iters; it++) {  // it == -1: warm-up
        if (flush) CU(cudaMemsetAsync(c->flush_buf, it & 0xff, c->flush_bytes, c->stream));
        CU(cudaEventRecord(e0, c->stream));
        for (const DevTensor* t : ts) {
            GemvParams p{};
            fill_seg(p.seg[0], *t, c->logits, nullptr, 0);
            p.n_seg = 1; p.K = (int)t->ne[0]; p.x = xin; p.epi = EPI_STORE;
            cudaError_t e = launch_gemv(c, p);
            if (e != cudaSuccess) { c->use_pdl = pdl; return fail(B200_ERR_OPERATION_FAILED, cudaGetErrorString(e)); }
        }
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 0) sum_ms += ms;
    }
    c->use_pdl = pdl;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(xin);
    if (avg_ms) *avg_ms = (float)(sum_ms / iters);
    if (bytes) *bytes = total;
    return B200_OK;
}

// ------------------------------------------------------------------ Backend per-op surface
// Host in / host out, like CudaBackend's per-op path (src/backend/cuda/mod.rs:229-835).
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 4) == cudaSuccess ? 0 : -1; }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
int op_ready() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B200_ERR_NOT_AVAILABLE, "cuda-b200: no CUDA device (this backend has no CPU fallback)");
    }
    return B200_OK;
}
int op_finish(const char* what) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B200_ERR_OPERATION_FAILED, std::string(what) + ": " + cudaGetErrorString(e));
    return B200_OK;
}
int elementwise(int op, const float* a, const float* b, float s, float* out, size_t n, const char* what) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!a || !out || (op <= EW_MUL && !b)) return fail(B200_ERR_INVALID_ARGUMENT, std::string(what) + ": null pointer");
    if (n == 0) return B200_OK;
    DevBuf da, db, dout;
    if (da.alloc(n * 4) || db.alloc(n * 4) || dout.alloc(n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, what);
    CU(cudaMemcpy(da.p, a, n * 4, cudaMemcpyHostToDevice));
    if (b) CU(cudaMemcpy(db.p, b, n * 4, cudaMemcpyHostToDevice));
    int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    elementwise_kernel<<<grid, 256>>>(op, da.as<float>(), db.as<float>(), s, dout.as<float>(), (long long)n);
    if ((rc = op_finish(what))) return rc;
    CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}
}  // namespace

extern "C" int b200_op_add(const float* a, const float* b, float* out, size_t n) { return elementwise(EW_ADD, a, b, 0, out, n, "add"); }
extern "C" int b200_op_mul(const float* a, const float* b, float* out, size_t n) { return elementwise(EW_MUL, a, b, 0, out, n, "mul"); }
extern "C" int b200_op_scale(const float* a, float s, float* out, size_t n) { return elementwise(EW_SCALE, a, nullptr, s, out, n, "scale"); }
extern "C" int b200_op_silu(const float* x, float* out, size_t n) { return elementwise(EW_SILU, x, nullptr, 0, out, n, "silu"); }
extern "C" int b200_op_gelu(const float* x, float* out, size_t n) { return elementwise(EW_GELU, x, nullptr, 0, out, n, "gelu"); }

extern "C" int b200_op_softmax(const float* x, float* out, size_t n) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!x || !out) return fail(B200_ERR_INVALID_ARGUMENT, "softmax: null pointer");
    if (n == 0) return B200_OK;
    DevBuf dx, dout;
    if (dx.alloc(n * 4) || dout.alloc(n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "softmax");
    CU(cudaMemcpy(dx.p, x, n * 4, cudaMemcpyHostToDevice));
    softmax_rows_kernel<<<1, 256>>>(dx.as<float>(), dout.as<float>(), (int)n);
    if ((rc = op_finish("softmax"))) return rc;
    CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_rms_norm(const float* x, const float* w, float eps, float* out, size_t n_rows, size_t hidden) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!x || !w || !out) return fail(B200_ERR_INVALID_ARGUMENT, "rms_norm: null pointer");
    if (n_rows == 0 || hidden == 0) return fail(B200_ERR_SHAPE_MISMATCH, "rms_norm: empty shape");
    size_t n = n_rows * hidden;
    DevBuf dx, dw, dout;
    if (dx.alloc(n * 4) || dw.alloc(hidden * 4) || dout.alloc(n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "rms_norm");
    CU(cudaMemcpy(dx.p, x, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.p, w, hidden * 4, cudaMemcpyHostToDevice));
    rms_norm_rows_kernel<<<(int)n_rows, 256>>>(dx.as<float>(), dw.as<float>(), eps, dout.as<float>(), (int)hidden);
    if ((rc = op_finish("rms_norm"))) return rc;
    CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_vec_mat(const float* a, const float* w, float* out, size_t k, size_t n) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!a || !w || !out) return fail(B200_ERR_INVALID_ARGUMENT, "vec_mat: null pointer");
    if (k == 0 || n == 0) return fail(B200_ERR_SHAPE_MISMATCH, "vec_mat: empty shape");
    DevBuf da, dw, dout;
    if (da.alloc(k * 4) || dw.alloc(k * n * 4) || dout.alloc(n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "vec_mat");
    CU(cudaMemcpy(da.p, a, k * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.p, w, k * n * 4, cudaMemcpyHostToDevice));
    vec_mat_f32_kernel<<<(int)((n * 32 + 255) / 256), 256>>>(da.as<float>(), dw.as<float>(), dout.as<float>(), (int)k, (int)n);
    if ((rc = op_finish("vec_mat"))) return rc;
    CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// Backend extension for prefill / batched decode: T input rows at once through the tcgen05 dequant-GEMM
// (csrc/gemm_umma.cuh).  out[t][j] = sum_k a[t][k] * deq(W)[j][k]: the same contraction as T calls of vec_mat_q, fp16
// operands with f32 accumulation (documented tolerance 2e-3 of the largest output).
// Backend::matmul / matvec / matvec_q (src/backend/mod.rs:90, 93, 107): required trait methods that the model code never calls
extern "C" int b200_op_matmul(const float* a, const float* b, float* out, size_t m, size_t k, size_t n) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!a || !b || !out) return fail(B200_ERR_INVALID_ARGUMENT, "matmul: null pointer");
    if (m == 0 || k == 0 || n == 0) return fail(B200_ERR_SHAPE_MISMATCH, "matmul: empty shape");
    DevBuf da, db, dout;
    if (da.alloc(m * k * 4) || db.alloc(k * n * 4) || dout.alloc(m * n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "matmul");
    CU(cudaMemcpy(da.p, a, m * k * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db.p, b, k * n * 4, cudaMemcpyHostToDevice));
    matmul_f32_kernel<<<(int)((m * n + 255) / 256), 256>>>(da.as<float>(), db.as<float>(), dout.as<float>(), (int)m, (int)k, (int)n);
    if ((rc = op_finish("matmul"))) return rc;
    CU(cudaMemcpy(out, dout.p, m * n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}
// matvec: a [m][k] row-major @ b [k] -> [m] (cpu/ops.rs:531-575): row i of a is k contiguous floats, exactly vec_mat's GGUF layout
extern "C" int b200_op_matvec(const float* a, const float* b, float* out, size_t m, size_t k) { return b200_op_vec_mat(b, a, out, k, m); }
// matvec_q: a = m rows of k / bs quantised blocks @ b [k] -> [m] (cpu/ops.rs:922-946): the block walk of vec_mat_q
extern "C" int b200_op_vec_mat_q(const float* a, const void* w, uint32_t ggml_type, float* out, size_t k, size_t n);
extern "C" int b200_op_matvec_q(const void* a, uint32_t ggml_type, const float* b, float* out, size_t m, size_t k) {
    return b200_op_vec_mat_q(b, a, ggml_type, out, k, m);
}

extern "C" int b200_op_mat_mat_q(const float* a, const void* w, uint32_t ggml_type, float* out, size_t t_rows, size_t k, size_t n) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!a || !w || !out) return fail(B200_ERR_INVALID_ARGUMENT, "mat_mat_q: null pointer");
    const int t = (int)ggml_type;
    if (!umma_type_ok(t)) return fail(B200_ERR_UNSUPPORTED_DTYPE, "mat_mat_q: ggml type " + std::to_string(ggml_type) + " (Q4_K, Q5_K, Q6_K, Q8_0 only)");
    const int be = type_block_elems(t), bb = type_block_bytes(t);
    if (t_rows == 0 || k == 0 || n == 0 || k % be || k % 64) return fail(B200_ERR_SHAPE_MISMATCH, "mat_mat_q: k must be a non-zero multiple of the block size and of 64");
    const size_t row_bytes = k / be * bb, wbytes = row_bytes * n;
    DevBuf da, dh, dw, dout;
    if (da.alloc(t_rows * k * 4) || dh.alloc(t_rows * k * 2) || dw.alloc(wbytes + 256) || dout.alloc(t_rows * n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "mat_mat_q");
    CU(cudaMemcpy(da.p, a, t_rows * k * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.p, w, wbytes, cudaMemcpyHostToDevice));
    umma_to_half_kernel<<<(int)std::min<size_t>((t_rows * k + 255) / 256, 148 * 8), 256>>>(da.as<float>(), dh.as<__half>(), (long long)(t_rows * k));
    UmmaParams p{};
    p.w = dw.as<uint8_t>(); p.row_bytes = (long long)row_bytes; p.type = t; p.n_rows = (int)n; p.K = (int)k;
    p.x = dh.as<__half>(); p.ldx = (int)k; p.T = (int)t_rows; p.y = dout.as<float>(); p.ldy = (int)n;
    if (!umma_eligible(p)) return fail(B200_ERR_SHAPE_MISMATCH, "mat_mat_q: rows of this type / length are not aligned for the tensor-core path");
    // raw-tile tensor map for the persistent kernel (what b200_ctx_finalize builds once per weight matrix)
    DevBuf dmap;
    if (gemm2_enabled() && gemm_encode_fn() && !(row_bytes & 15) && k % 256 == 0 && !dmap.alloc(sizeof(CUtensorMap))) {
        const int pitch = stream_pitch(t, 1);
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 4), (cuuint64_t)n};
        const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)(pitch / 4), (cuuint32_t)kUmmaM};
        const cuuint32_t estr[2] = {1, 1};
        if (gemm_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dw.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
            CU(cudaMemcpy(dmap.p, &tm, sizeof tm, cudaMemcpyHostToDevice));
            p.tmap = dmap.p; p.raw_pitch = pitch; p.raw_bytes = 256 / be * bb;
        }
    }
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    CU(gemm_dispatch(p, n_sm, 0));
    if ((rc = op_finish("mat_mat_q"))) return rc;
    CU(cudaMemcpy(out, dout.p, t_rows * n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_vec_mat_q(const float* a, const void* w, uint32_t ggml_type, float* out, size_t k, size_t n) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!a || !w || !out) return fail(B200_ERR_INVALID_ARGUMENT, "vec_mat_q: null pointer");
    const int t = (int)ggml_type;
    if (!type_supported(t)) return fail(B200_ERR_UNSUPPORTED_DTYPE, "vec_mat_q: unsupported ggml type " + std::to_string(ggml_type));
    const int be = type_block_elems(t), bb = type_block_bytes(t);
    if (k == 0 || n == 0 || k % be || k % 32) return fail(B200_ERR_SHAPE_MISMATCH, "vec_mat_q: k must be a non-zero multiple of the block size and of 32");
    const size_t row_bytes = k / be * bb, wbytes = row_bytes * n;
    DevBuf da, dw, dout;
    if (da.alloc(k * 4) || dw.alloc(wbytes + 256) || dout.alloc(n * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "vec_mat_q");
    CU(cudaMemcpy(da.p, a, k * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.p, w, wbytes, cudaMemcpyHostToDevice));
    // tensor-pipe kernel when the type/shape is eligible (the path the model uses), else the CUDA-core kernel
    {
        cudaDeviceProp prop{};
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaGetDeviceProperties(&prop, dev));
        MParams m{};
        m.seg[0].w = dw.as<uint8_t>(); m.seg[0].out = dout.as<float>(); m.seg[0].row_bytes = (long long)row_bytes;
        m.seg[0].type = t; m.seg[0].n_rows = (int)n;
        m.n_seg = 1; m.K = (int)k; m.x = da.as<float>(); m.epi = ME_STORE;
        MPlan plan;
        const size_t lim = (size_t)prop.sharedMemPerBlockOptin - 8192;
        DevBuf dpart, dtick;
        if (env_int("B200_GEMV_MMA", 1) && mma_plan(m, prop.multiProcessorCount, 16, 3, lim, plan)) {
            const size_t tiles = (n + 15) / 16;
            if (dpart.alloc((size_t)plan.grid * 2 * 64 * 4) || dtick.alloc(tiles * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "vec_mat_q");
            CU(cudaMemset(dtick.p, 0, tiles * 4));
            m.part = dpart.as<float>(); m.tickets = dtick.as<unsigned int>();
            CU(mma_set_smem_limit((int)lim));
            mma_kernel_for(plan.stages)<<<plan.grid, plan.warps * 32, plan.smem>>>(m);
            if ((rc = op_finish("vec_mat_q"))) return rc;
            CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
            return B200_OK;
        }
    }
    CU(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    GemvParams p{};
    p.seg[0].w = dw.as<uint8_t>(); p.seg[0].out = dout.as<float>(); p.seg[0].row_bytes = (long long)row_bytes;
    p.seg[0].type = t; p.seg[0].n_rows = (int)n;
    p.n_seg = 1; p.K = (int)k; p.x = da.as<float>(); p.epi = EPI_STORE;
    int n_tasks = (int)((n + kGemvR - 1) / kGemvR);
    int grid = std::max(1, std::min((n_tasks + kGemvWarps - 1) / kGemvWarps, 296));
    size_t smem = (size_t)xpad_floats((int)k) * sizeof(float);
    if (smem > 160 * 1024) return fail(B200_ERR_UNSUPPORTED, "vec_mat_q: k too large for the shared-memory x staging");
    gemv_kernel<<<grid, kGemvThreads, smem>>>(p);
    if ((rc = op_finish("vec_mat_q"))) return rc;
    CU(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_dequantize(const void* src, uint32_t ggml_type, float* out, size_t n_elems) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!src || !out) return fail(B200_ERR_INVALID_ARGUMENT, "dequantize: null pointer");
    const int t = (int)ggml_type;
    if (!type_supported(t)) return fail(B200_ERR_UNSUPPORTED_DTYPE, "dequantize: unsupported ggml type " + std::to_string(ggml_type));
    const int be = type_block_elems(t), bb = type_block_bytes(t);
    if (n_elems % be) return fail(B200_ERR_SHAPE_MISMATCH, "dequantize: element count not a multiple of the block size");
    if (n_elems == 0) return B200_OK;
    const size_t nbytes = n_elems / be * bb;
    DevBuf ds, dout;
    if (ds.alloc(nbytes) || dout.alloc(n_elems * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "dequantize");
    CU(cudaMemcpy(ds.p, src, nbytes, cudaMemcpyHostToDevice));
    int grid = (int)std::min<size_t>((n_elems + 255) / 256, 148 * 16);
    dequantize_kernel<<<grid, 256>>>(t, ds.as<uint8_t>(), (long long)n_elems, dout.as<float>());
    if ((rc = op_finish("dequantize"))) return rc;
    CU(cudaMemcpy(out, dout.p, n_elems * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_rope(float* q, float* k, int n_heads, int n_kv_heads, int head_dim, int pos, float freq_base,
                            float freq_scale, int use_neox) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!q || !k) return fail(B200_ERR_INVALID_ARGUMENT, "rope: null pointer");
    if (n_heads <= 0 || n_kv_heads <= 0 || head_dim <= 0 || (head_dim & 1) || pos < 0)
        return fail(B200_ERR_INVALID_ARGUMENT, "RoPE requires 3D tensors [num_heads, seq_len, head_dim] with even head_dim");
    if (freq_scale == 0.0f) freq_scale = 1.0f;
    std::vector<float> f(head_dim / 2);
    for (int i = 0; i < head_dim / 2; i++) f[i] = 1.0f / powf(freq_base, (float)(2 * i) / (float)head_dim);
    const size_t nq = (size_t)n_heads * head_dim, nk = (size_t)n_kv_heads * head_dim;
    DevBuf dq, dk, df;
    if (dq.alloc(nq * 4) || dk.alloc(nk * 4) || df.alloc(f.size() * 4)) return fail(B200_ERR_ALLOCATION_FAILED, "rope");
    CU(cudaMemcpy(dq.p, q, nq * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dk.p, k, nk * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(df.p, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    int work = (n_heads + n_kv_heads) * head_dim / 2;
    rope_inplace_kernel<<<(work + 255) / 256, 256>>>(dq.as<float>(), dk.as<float>(), df.as<float>(), n_heads, n_kv_heads,
                                                     head_dim, pos, freq_scale, use_neox);
    if ((rc = op_finish("rope"))) return rc;
    CU(cudaMemcpy(q, dq.p, nq * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(k, dk.p, nk * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_attention_cached(const float* q, const float* k_cache, const float* v_cache, float* out,
                                        int n_heads, int n_kv_heads, int head_dim, int max_seq, float scale, int kv_len) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!q || !k_cache || !v_cache || !out) return fail(B200_ERR_INVALID_ARGUMENT, "attention_cached: null pointer");
    if (n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads || kv_len <= 0 || kv_len > max_seq)
        return fail(B200_ERR_INVALID_ARGUMENT, "attention_cached: bad head counts or kv_len");
    const int G = n_heads / n_kv_heads;
    if ((head_dim != 64 && head_dim != 128) || G > 8)
        return fail(B200_ERR_UNSUPPORTED, "attention_cached: head_dim must be 64 or 128 and at most 8 query heads per kv head");
    const size_t nq = (size_t)n_heads * head_dim, nc = (size_t)n_kv_heads * max_seq * head_dim;
    const int n_splits = std::max(1, std::min(32, (kv_len + 63) / 64));
    DevBuf dq, dk, dv, dout, dpart, dtick;
    if (dq.alloc(nq * 4) || dk.alloc(nc * 4) || dv.alloc(nc * 4) || dout.alloc(nq * 4) ||
        dpart.alloc((size_t)n_kv_heads * n_splits * G * (head_dim + 2) * 4) || dtick.alloc(n_kv_heads * 4))
        return fail(B200_ERR_ALLOCATION_FAILED, "attention_cached");
    CU(cudaMemcpy(dq.p, q, nq * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dk.p, k_cache, nc * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dv.p, v_cache, nc * 4, cudaMemcpyHostToDevice));
    CU(cudaMemset(dtick.p, 0, n_kv_heads * 4));
    AttnParams ap{};
    ap.q = dq.as<float>(); ap.k_cache = dk.as<float>(); ap.v_cache = dv.as<float>(); ap.out = dout.as<float>();
    ap.part = dpart.as<float>(); ap.tickets = dtick.as<unsigned int>(); ap.pos = nullptr; ap.kv_len_fixed = kv_len;
    ap.n_kv = n_kv_heads; ap.G = G; ap.max_seq = max_seq; ap.n_splits = n_splits; ap.scale = scale;
    dim3 grid(n_kv_heads, n_splits);
    {
        cudaDeviceProp prop{};
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaGetDeviceProperties(&prop, dev));
        const int lim_attn = (int)prop.sharedMemPerBlockOptin - 1024;
        CU(cudaFuncSetAttribute(attn_decode_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
        CU(cudaFuncSetAttribute(attn_decode_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
        CU(cudaFuncSetAttribute(attn_decode_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
        CU(cudaFuncSetAttribute(attn_decode_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim_attn));
    }
    if (head_dim == 128) {
        if (G <= 4) attn_decode_kernel<128, 4><<<grid, kAttnThreads, attn_smem_bytes(128, 4, ap.n_splits, ap.G)>>>(ap);
        else attn_decode_kernel<128, 8><<<grid, kAttnThreads, attn_smem_bytes(128, 8, ap.n_splits, ap.G)>>>(ap);
    } else {
        if (G <= 4) attn_decode_kernel<64, 4><<<grid, kAttnThreads, attn_smem_bytes(64, 4, ap.n_splits, ap.G)>>>(ap);
        else attn_decode_kernel<64, 8><<<grid, kAttnThreads, attn_smem_bytes(64, 8, ap.n_splits, ap.G)>>>(ap);
    }
    if ((rc = op_finish("attention_cached"))) return rc;
    CU(cudaMemcpy(out, dout.p, nq * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_op_attention(const float* q, const float* k, const float* v, float* out, int n_heads, int n_kv_heads,
                                 int seq_len, int head_dim, float scale) {
    int rc;
    if ((rc = op_ready())) return rc;
    if (!q || !k || !v || !out) return fail(B200_ERR_INVALID_ARGUMENT, "attention: null pointer");
    if (n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads || seq_len <= 0 || head_dim <= 0 || head_dim > 256)
        return fail(B200_ERR_INVALID_ARGUMENT, "Attention requires 3D tensors with head_dim <= 256");
    const size_t nq = (size_t)n_heads * seq_len * head_dim, nk = (size_t)n_kv_heads * seq_len * head_dim;
    DevBuf dq, dk, dv, dout;
    if (dq.alloc(nq * 4) || dk.alloc(nk * 4) || dv.alloc(nk * 4) || dout.alloc(nq * 4))
        return fail(B200_ERR_ALLOCATION_FAILED, "attention");
    CU(cudaMemcpy(dq.p, q, nq * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dk.p, k, nk * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dv.p, v, nk * 4, cudaMemcpyHostToDevice));
    int warps = n_heads * seq_len;
    attention_full_kernel<<<(warps * 32 + 255) / 256, 256>>>(dq.as<float>(), dk.as<float>(), dv.as<float>(), dout.as<float>(),
                                                            n_heads, n_kv_heads, seq_len, seq_len, head_dim, scale);
    if ((rc = op_finish("attention"))) return rc;
    CU(cudaMemcpy(out, dout.p, nq * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// ------------------------------------------------------------------ GGUF -> HBM direct load (gguf_load.cuh)
extern "C" int b200_gguf_open(const char* path, b200_gguf** out) {
    if (!path || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_gguf_open: null argument");
    *out = nullptr;
    b200_gguf* g = new b200_gguf();
    g->fd = open(path, O_RDONLY);
    struct stat st;
    if (g->fd < 0 || fstat(g->fd, &st) != 0) {
        gguf_release(g);
        return fail(B200_ERR_INVALID_ARGUMENT, std::string("b200_gguf_open: cannot open ") + path);
    }
    g->size = (size_t)st.st_size;
    if (g->size < 8) {
        gguf_release(g);
        return fail(B200_ERR_INVALID_ARGUMENT, std::string(path) + ": not a GGUF file (shorter than its header)");
    }
    void* m = mmap(nullptr, g->size, PROT_READ, MAP_PRIVATE, g->fd, 0);
    if (m == MAP_FAILED) {
        g->map = nullptr;
        gguf_release(g);
        return fail(B200_ERR_ALLOCATION_FAILED, std::string("b200_gguf_open: mmap failed for ") + path);
    }
    g->map = (const uint8_t*)m;
    madvise(m, g->size, MADV_SEQUENTIAL);
    GgufCursor c{g->map, g->size};
    const uint32_t magic = c.get<uint32_t>();
    if (magic != 0x46554747u) {   // GgufError::InvalidMagic (reader.rs:35)
        gguf_release(g);
        char b[64];
        snprintf(b, sizeof b, "invalid GGUF magic 0x%08x", magic);
        return fail(B200_ERR_INVALID_ARGUMENT, b);
    }
    g->version = c.get<uint32_t>();
    if (g->version < 1 || g->version > 3) {   // GgufError::UnsupportedVersion (reader.rs:41)
        const uint32_t v = g->version;
        gguf_release(g);
        return fail(B200_ERR_UNSUPPORTED, "unsupported GGUF version " + std::to_string(v));
    }
    const uint64_t n_tensors = c.len(g->version), n_kv = c.len(g->version);
    if (!c.ok || n_tensors > g->size || n_kv > g->size) {
        gguf_release(g);
        return fail(B200_ERR_INVALID_ARGUMENT, "GGUF header truncated or counts out of range");
    }
    for (uint64_t i = 0; i < n_kv; i++) {
        std::string key = c.str(g->version);
        const uint32_t type = c.get<uint32_t>();
        GgufValue v;
        if (!c.ok || !gguf_read_value(c, g->version, type, v, 0)) {
            const bool trunc = !c.ok;
            gguf_release(g);
            return fail(B200_ERR_INVALID_ARGUMENT, trunc ? "GGUF metadata truncated (key '" + key + "')"
                                                         : "invalid GGUF metadata value type " + std::to_string(type) + " (key '" + key + "')");
        }
        g->kv[key] = std::move(v);
    }
    g->tensors.reserve((size_t)n_tensors);
    for (uint64_t i = 0; i < n_tensors; i++) {
        GgufTensorInfo t;
        t.name = c.str(g->version);
        const uint32_t nd = c.get<uint32_t>();
        if (!c.ok || nd < 1 || nd > 4) {
            gguf_release(g);
            return fail(B200_ERR_INVALID_ARGUMENT, "GGUF tensor info truncated or with a bad rank (tensor " + std::to_string(i) + ")");
        }
        t.n_dims = (int)nd;
        for (uint32_t k = 0; k < nd; k++) t.ne[k] = c.len(g->version);
        t.type = c.get<uint32_t>();
        t.offset = c.get<uint64_t>();
        if (!c.ok) {
            gguf_release(g);
            return fail(B200_ERR_INVALID_ARGUMENT, "GGUF tensor infos truncated");
        }
        const int be = type_block_elems((int)t.type), bb = type_block_bytes((int)t.type);
        if (be) {
            uint64_t numel = 1;
            bool overflow = false;
            for (int k = 0; k < t.n_dims; k++) overflow |= __builtin_mul_overflow(numel, t.ne[k], &numel);
            if (overflow || numel / be > ((uint64_t)1 << 50)) {   // (the byte size below must not wrap; real bounds: b200_gguf_tensor_info)
                const std::string nm = t.name;
                gguf_release(g);
                return fail(B200_ERR_INVALID_ARGUMENT, nm + ": tensor shape overflows (corrupt GGUF?)");
            }
            t.nbytes = (size_t)(numel / be * bb);   // TensorInfo::data_size (types.rs:45-50)
        }
        g->tensors.push_back(std::move(t));
    }
    {   // general.alignment: Uint32 or Uint64, default 32 (reader.rs:84-96)
        auto it = g->kv.find("general.alignment");
        if (it != g->kv.end() && (it->second.type == GV_U32 || it->second.type == GV_U64) && it->second.u > 0) g->alignment = it->second.u;
    }
    g->data_offset = ((uint64_t)c.pos + g->alignment - 1) / g->alignment * g->alignment;
    {
        auto it = g->kv.find("general.architecture");
        if (it != g->kv.end() && it->second.type == GV_STRING) g->arch = it->second.s;
    }
    *out = g;
    return B200_OK;
}

extern "C" void b200_gguf_close(b200_gguf* g) { gguf_release(g); }

extern "C" int b200_gguf_info(b200_gguf* g, uint32_t* version, uint64_t* n_tensors, uint64_t* n_metadata, uint64_t* alignment,
                              uint64_t* data_offset, uint64_t* file_bytes) {
    if (!g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_gguf_info: null file");
    if (version) *version = g->version;
    if (n_tensors) *n_tensors = g->tensors.size();
    if (n_metadata) *n_metadata = g->kv.size();
    if (alignment) *alignment = g->alignment;
    if (data_offset) *data_offset = g->data_offset;
    if (file_bytes) *file_bytes = g->size;
    return B200_OK;
}

extern "C" int b200_gguf_architecture(b200_gguf* g, char* out, size_t cap) {
    if (!g || !out || cap == 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_gguf_architecture: bad argument");
    if (g->arch.empty()) return fail(B200_ERR_INVALID_ARGUMENT, "missing metadata general.architecture");
    snprintf(out, cap, "%s", g->arch.c_str());
    return B200_OK;
}

// (name, type, ne, n_dims, pointer into the mapping, bytes) of tensor i; *data is NULL (and *nbytes 0) for a type the engine does not
// implement.  Fails if the tensor's bytes lie outside the file (GgufFile::tensor_data returns None, mod.rs:34-42).
extern "C" int b200_gguf_tensor_info(b200_gguf* g, uint64_t i, const char** name, uint32_t* ggml_type, uint64_t* ne4, int* n_dims,
                                     const void** data, size_t* nbytes) {
    if (!g || i >= g->tensors.size()) return fail(B200_ERR_INVALID_ARGUMENT, "b200_gguf_tensor_info: bad argument");
    const GgufTensorInfo& t = g->tensors[(size_t)i];
    if (name) *name = t.name.c_str();
    if (ggml_type) *ggml_type = t.type;
    if (ne4) for (int k = 0; k < 4; k++) ne4[k] = t.ne[k];
    if (n_dims) *n_dims = t.n_dims;
    if (nbytes) *nbytes = t.nbytes;
    if (data) {
        *data = nullptr;
        if (t.nbytes) {
            const uint64_t start = g->data_offset + t.offset;
            if (t.offset > g->size || start > g->size || t.nbytes > g->size - start)
                return fail(B200_ERR_INVALID_ARGUMENT, t.name + ": tensor data lies outside the file (truncated GGUF?)");
            *data = g->map + start;
        }
    }
    return B200_OK;
}

// ModelLoader::parse_config (src/model/loader.rs:62-300), the keys GpuOnlyInference needs, same defaults and fallbacks.
extern "C" int b200_gguf_model_desc(b200_gguf* g, int max_seq_len, int max_batch, b200_model_desc* out) {
    if (!g || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_gguf_model_desc: null argument");
    if (g->arch.empty()) return fail(B200_ERR_INVALID_ARGUMENT, "missing metadata general.architecture");
    const std::string& a = g->arch;
    // gemma (GELU), phi / gptneox (LayerNorm, fused QKV) ... share tensor names with llama but not its arithmetic: refuse them
    if (a != "llama" && a != "mistral" && a != "mixtral" && a != "qwen2" && a != "qwen2moe")
        return fail(B200_ERR_UNSUPPORTED, "unsupported architecture '" + a + "': cuda-b200 implements llama, mistral, mixtral, qwen2, qwen2moe (RMSNorm + SwiGLU + RoPE)");
    auto need_u32 = [&](const char* key, uint32_t& v) { return gguf_get_u32(g, a + "." + key, v); };
    b200_model_desc d{};
    uint32_t u = 0;
    if (!need_u32("embedding_length", u)) return fail(B200_ERR_INVALID_ARGUMENT, "missing metadata " + a + ".embedding_length");
    d.hidden = (int32_t)u;
    if (!need_u32("block_count", u)) return fail(B200_ERR_INVALID_ARGUMENT, "missing metadata " + a + ".block_count");
    d.n_layers = (int32_t)u;
    if (!need_u32("attention.head_count", u) || u == 0) return fail(B200_ERR_INVALID_ARGUMENT, "missing metadata " + a + ".attention.head_count");
    d.n_heads = (int32_t)u;
    d.n_kv_heads = need_u32("attention.head_count_kv", u) ? (int32_t)u : d.n_heads;
    d.head_dim = need_u32("attention.key_length", u) ? (int32_t)u : d.hidden / d.n_heads;
    d.ffn = need_u32("feed_forward_length", u) ? (int32_t)u : d.hidden * 4 * 2 / 3;
    const int file_ctx = need_u32("context_length", u) ? (int)u : 2048;
    d.max_seq_len = max_seq_len > 0 ? std::min(max_seq_len, file_ctx) : file_ctx;
    float f = 0.0f;
    d.norm_eps = gguf_get_f32(g, a + ".attention.layer_norm_rms_epsilon", f) ? f : gguf_get_f32(g, a + ".attention.layer_norm_epsilon", f) ? f : 1e-5f;
    d.rope_base = gguf_get_f32(g, a + ".rope.freq_base", f) ? f : 10000.0f;
    d.rope_scale = gguf_get_f32(g, a + ".rope.scale_linear", f) ? f : 1.0f;
    d.rope_neox = (a == "qwen2" || a == "qwen2moe") ? 1 : 0;   // loader.rs:145-162
    d.n_experts = need_u32("expert_count", u) ? (int32_t)u : 0;
    d.n_experts_used = need_u32("expert_used_count", u) ? (int32_t)u : 0;
    d.expert_ffn = need_u32("expert_feed_forward_length", u) ? (int32_t)u : 0;
    if (need_u32("rope.dimension_count", u) && (int32_t)u != d.head_dim)
        return fail(B200_ERR_UNSUPPORTED, "partial RoPE (rope.dimension_count " + std::to_string(u) + " != head_dim " + std::to_string(d.head_dim) + ") is not implemented");
    // vocab: {arch}.vocab_size, tokenizer.ggml.vocab_size, the token array's length, the embedding's rows, 32000 (loader.rs:77-98)
    const GgufTensorInfo* emb = gguf_find_tensor(g, "token_embd.weight");
    if (need_u32("vocab_size", u) || gguf_get_u32(g, "tokenizer.ggml.vocab_size", u)) d.vocab = (int32_t)u;
    else {
        auto it = g->kv.find("tokenizer.ggml.tokens");
        if (it != g->kv.end() && it->second.type == GV_ARRAY) d.vocab = (int32_t)it->second.arr_len;
        else if (emb && emb->n_dims == 2) d.vocab = (int32_t)emb->ne[1];
        else d.vocab = 32000;
    }
    d.tied_output = gguf_find_tensor(g, "output.weight") ? 0 : 1;   // loader.rs:349-355
    if (d.n_experts > 0 && d.expert_ffn == 0) {
        const GgufTensorInfo* ge = gguf_find_tensor(g, "blk.0.ffn_gate_exps.weight");
        if (ge && ge->n_dims == 3) d.expert_ffn = (int32_t)ge->ne[1];
    }
    d.max_batch = std::max(1, max_batch);
    *out = d;
    return B200_OK;
}

static cudaError_t stage_begin(b200_ctx* c) {
    b200_ctx::LoadStage& s = c->stage;
    cudaError_t e;
    s.chunk = (size_t)std::max(1, env_int("B200_LOAD_CHUNK_MB", 32)) << 20;
    if (env_int("B200_LOAD_CHUNK_KB", 0) > 0) s.chunk = (size_t)env_int("B200_LOAD_CHUNK_KB", 0) << 10;   // (tests: many chunks per tensor)
    // measured on the B200 box (16 host cores, Llama-3-8B Q4_K_M, 4.9 GB from the page cache): 1 thread 5.8 GB/s, 4: 9-14, 8: 15.8
    s.threads = std::max(1, env_int("B200_LOAD_THREADS", (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()))));
    s.next = 0;
    s.bytes = 0;
    if ((e = cudaStreamCreate(&s.st)) != cudaSuccess) return e;   // blocking stream: ordered with the legacy-stream memsets
    for (int b = 0; b < 2; b++) {
        if ((e = cudaHostAlloc((void**)&s.pin[b], s.chunk, cudaHostAllocDefault)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&s.ev[b], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    s.on = true;
    return cudaSuccess;
}
static cudaError_t stage_end(b200_ctx* c) {
    b200_ctx::LoadStage& s = c->stage;
    s.on = false;
    cudaError_t e = s.st ? cudaStreamSynchronize(s.st) : cudaSuccess;
    for (int b = 0; b < 2; b++) {
        if (s.ev[b]) cudaEventDestroy(s.ev[b]);
        if (s.pin[b]) cudaFreeHost(s.pin[b]);
        s.ev[b] = nullptr;
        s.pin[b] = nullptr;
    }
    if (s.st) cudaStreamDestroy(s.st);
    s.st = nullptr;
    return e;
}

// Every tensor of the file that the engine has a slot for goes to HBM (this rank's shard under tensor / expert parallelism);
// tensors it does not use (rope_freqs.weight, tokenizer tables ...) are skipped and counted.  B200_LOAD_STAGED=0 falls back to
// cudaMemcpy straight from the mapping (pageable: the driver stages it internally, one thread) for comparison.
extern "C" int b200_ctx_load_gguf(b200_ctx* c, b200_gguf* g, b200_load_stats* stats) {
    if (!c || !g) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_load_gguf: null argument");
    if (c->finalized) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_load_gguf: context already finalized");
    CU(cudaSetDevice(c->par.device));
    const auto t0 = std::chrono::steady_clock::now();
    const bool staged = env_int("B200_LOAD_STAGED", 1) != 0;
    c->stage.bytes = 0;
    if (staged) {
        cudaError_t e = stage_begin(c);
        if (e != cudaSuccess) {
            stage_end(c);
            return fail(B200_ERR_ALLOCATION_FAILED, std::string("b200_ctx_load_gguf: staging buffers: ") + cudaGetErrorString(e));
        }
    }
    if (!c->arena && env_int("B200_LOAD_ARENA", 1)) {   // one zeroed allocation for all the tensors (upper bound: full tensors + padding)
        const size_t P = (size_t)std::max(1, c->par.world_size);
        size_t total = 0;
        for (const GgufTensorInfo& t : g->tensors) {
            if (!slot_for_name(c, t.name) || !t.nbytes) continue;
            size_t b = t.nbytes;   // this rank's bytes of the tensor, as b200_ctx_upload_tensor will cut it
            if (P > 1 && c->ep) {
                if (t.name.find("_exps.weight") != std::string::npos) b = t.nbytes / P;
            } else if (P > 1) {
                const int kind = tp_shard_kind(t.name.c_str());
                const size_t row = (size_t)(t.ne[0] / type_block_elems((int)t.type) * type_block_bytes((int)t.type));
                if (kind == 1) b = t.nbytes / P;
                else if (kind == 2 && row) b = ((row / P + 15) & ~(size_t)15) * (t.nbytes / row);   // 16-byte row pitch
            }
            total += (b + 256 + 255) & ~(size_t)255;
        }
        if (total && cudaMalloc((void**)&c->arena, total) == cudaSuccess) {
            c->arena_size = total;
            c->arena_off = 0;
            CU(cudaMemset(c->arena, 0, total));
        } else {
            cudaGetLastError();
            c->arena = nullptr;   // (falls back to one allocation per tensor)
        }
    }
    uint32_t loaded = 0, skipped = 0;
    uint64_t bytes = 0;
    int rc = B200_OK;
    for (uint64_t i = 0; i < g->tensors.size() && rc == B200_OK; i++) {
        const GgufTensorInfo& t = g->tensors[(size_t)i];
        if (!slot_for_name(c, t.name)) { skipped++; continue; }
        const void* data = nullptr;
        size_t nbytes = 0;
        if ((rc = b200_gguf_tensor_info(g, i, nullptr, nullptr, nullptr, nullptr, &data, &nbytes))) break;
        if (!data) { rc = fail(B200_ERR_UNSUPPORTED_DTYPE, t.name + ": unsupported ggml type " + std::to_string(t.type)); break; }
        if (t.n_dims > 3) { rc = fail(B200_ERR_SHAPE_MISMATCH, t.name + ": 1..3 dims expected"); break; }
        rc = b200_ctx_upload_tensor(c, t.name.c_str(), t.type, t.ne, t.n_dims, data, nbytes);
        loaded++;
        bytes += nbytes;
    }
    const uint64_t staged_bytes = c->stage.bytes;
    if (staged) {
        cudaError_t e = stage_end(c);
        if (rc == B200_OK && e != cudaSuccess) rc = fail(B200_ERR_OPERATION_FAILED, std::string("b200_ctx_load_gguf: ") + cudaGetErrorString(e));
    } else {
        cudaDeviceSynchronize();
    }
    if (stats) {
        stats->file_bytes = g->size;
        stats->tensor_bytes = bytes;
        stats->device_bytes = staged_bytes;
        stats->tensors_loaded = loaded;
        stats->tensors_skipped = skipped;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return rc;
}

// GpuOnlyInference::from_model for a file path: open + describe + create + load + finalize in one call.
extern "C" int b200_ctx_create_from_gguf(const char* path, const b200_parallel_desc* par, int max_seq_len, int max_batch, int kv_format,
                                         b200_ctx** out, b200_load_stats* stats) {
    if (!path || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_create_from_gguf: null argument");
    *out = nullptr;
    b200_gguf* g = nullptr;
    int rc = b200_gguf_open(path, &g);
    if (rc) return rc;
    b200_model_desc d{};
    b200_ctx* c = nullptr;
    if ((rc = b200_gguf_model_desc(g, max_seq_len, max_batch, &d)) == B200_OK && (rc = b200_ctx_create(&d, par, &c)) == B200_OK) {
        if (kv_format != 0) rc = b200_ctx_set_kv_format(c, kv_format);
        if (rc == B200_OK) rc = b200_ctx_load_gguf(c, g, stats);
        if (rc == B200_OK) rc = b200_ctx_finalize(c);
    }
    const std::string msg = g_err;   // (destroy / close must not lose the message)
    b200_gguf_close(g);
    if (rc != B200_OK) {
        if (c) b200_ctx_destroy(c);
        g_err = msg;
        return rc;
    }
    *out = c;
    return B200_OK;
}

extern "C" int b200_group_load_gguf(b200_group* g, b200_gguf* file, b200_load_stats* stats) {
    if (!g || !file) return fail(B200_ERR_INVALID_ARGUMENT, "b200_group_load_gguf: null argument");
    return group_run(g, [=](int r) { return b200_ctx_load_gguf(g->ctx[r], file, r == 0 ? stats : nullptr); });
}

// ------------------------------------------------------------------ batched decode with the pick on the device + continuous batching
static bool batch_gemm_eligible(b200_ctx* c, int n) {
    if (n < c->batch_gemm_min || !prefill_gemm_ok(c)) return false;
    const DevTensor& head = c->output.present() ? c->output : c->token_embd;
    UmmaParams hp{};
    hp.w = head.d; hp.row_bytes = head.row_bytes; hp.type = head.type; hp.n_rows = c->d.vocab; hp.K = c->d.hidden; hp.T = 1;
    hp.x = reinterpret_cast<const __half*>(c->xa); hp.ldx = c->d.hidden;
    return umma_eligible(hp) && n <= std::min(prefill_chunk(), c->d.max_seq_len);
}

// b200_decode_batch with the greedy pick (last maximum wins, src/main.rs:1816-1821) on the device: n token ids come back instead
// of n x vocab logits (16 MB per step at batch 32 on Llama-3).  Same two paths as b200_decode_batch.
extern "C" int b200_decode_batch_greedy(b200_ctx* c, const int* seqs, const uint32_t* tokens, int n, uint32_t* next_out) {
    if (!c || !seqs || !tokens || !next_out || n <= 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_batch_greedy: bad argument");
    if (c->par.world_size > 1) return fail(B200_ERR_UNSUPPORTED, "b200_decode_batch_greedy: not available under tensor parallelism");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (seqs[i] == seqs[j]) return fail(B200_ERR_INVALID_ARGUMENT, "b200_decode_batch_greedy: duplicate sequence slot");
    int rc;
    for (int i = 0; i < n; i++) {
        if ((rc = check_slot(c, seqs[i], "b200_decode_batch_greedy"))) return rc;
        if ((rc = check_token(c, seqs[i], tokens[i], "b200_decode_batch_greedy"))) return rc;
        if ((rc = flush_pending(c, seqs[i]))) return rc;
    }
    if (batch_gemm_eligible(c, n)) {
        CU(cudaSetDevice(c->par.device));
        if (c->pf_argmax_rows < n) {
            cudaFree(c->pf_argmax);
            c->pf_argmax = nullptr;
            CU_ALLOC(cudaMalloc((void**)&c->pf_argmax, (size_t)n * sizeof(int)));
            c->pf_argmax_rows = n;
        }
        if ((rc = prefill_gemm(c, seqs[0], tokens, n, false, seqs))) return rc;
        argmax_rows_kernel<<<n, 1024, 0, c->stream>>>(c->pf_logits, c->d.vocab, c->pf_argmax);
        c->launches++;
        CU(cudaMemcpyAsync(next_out, c->pf_argmax, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        WATCHDOG_FETCH(c);
        CU(cudaStreamSynchronize(c->stream));
        if ((rc = watchdog_check(c, "b200_decode_batch_greedy"))) return rc;
        for (int i = 0; i < n; i++) c->slots[seqs[i]].host_pos++;
        return B200_OK;
    }
    for (int i = 0; i < n; i++)
        if ((rc = b200_decode_greedy(c, seqs[i], tokens[i], 1, next_out + i, nullptr))) return rc;
    return B200_OK;
}

// Continuous batching: the peer of BatchedEngine's background loop (src/engine_batched.rs:199-411) below the tokenizer and the
// channels.  submit = BatchedEngine::submit (:167-192, "queue full" past max_queue_depth) + the loop's drain step (:217-239: active
// while there is room, else pending); step = one iteration of the loop: step_sequence's rules for every active sequence (:366-411:
// finished if its last token is EOS or generated >= max_tokens, the whole prompt on the first step, then one token per step; EOS
// ends the sequence without a Token event; Done carries MaxTokens iff generated >= max_tokens) and then the LIFO promotion of
// pending requests (:291-304, Vec::pop).  The reference steps the sequences one by one through Model::forward; here every
// sequence owns a KV slot of the context, the prompts go through b200_prefill and ALL decoding sequences share one pass over the
// weights (b200_decode_batch_greedy).  Sampling is the greedy rule on the device.
struct BatchSeq {
    uint64_t id;
    std::vector<uint32_t> tokens;
    int prompt_len, generated, max_tokens, slot;
    bool started;
};
struct BatchReq {
    uint64_t id;
    std::vector<uint32_t> tokens;
    int max_tokens;
};
struct b200_batch {
    b200_ctx* c = nullptr;
    b200_batch_config cfg{};
    std::vector<BatchSeq> active;
    std::vector<BatchReq> pending;
    std::vector<int> free_slots;
    std::vector<b200_batch_event> events;   // not yet handed to the caller
    uint64_t next_id = 1;
    int queue_count = 0;                    // active + pending, as BatchedEngine::queue_count
    uint64_t steps = 0, batched_rows = 0;
    std::string last_error;                 // message of the last Error event
};

static void batch_event(b200_batch* b, uint64_t id, int kind, uint32_t token, int reason, int prompt_tokens, int completion_tokens) {
    b200_batch_event e{};
    e.request_id = id; e.kind = kind; e.token = token; e.reason = reason; e.prompt_tokens = prompt_tokens; e.completion_tokens = completion_tokens;
    b->events.push_back(e);
}
// create_active_sequence (engine_batched.rs:322-356)
static void batch_activate(b200_batch* b, BatchReq&& r) {
    if (r.tokens.empty()) {   // "empty prompt"
        batch_event(b, r.id, B200_BATCH_ERROR, 0, B200_FINISH_ERROR, 0, 0);
        b->last_error = "empty prompt";
        b->queue_count = std::max(0, b->queue_count - 1);
        return;
    }
    const int max_seq = std::min(b->cfg.max_seq_len, b->c->d.max_seq_len);
    const size_t keep = std::min(r.tokens.size(), (size_t)std::max(0, max_seq - 1));
    BatchSeq s;
    s.id = r.id;
    s.tokens.assign(r.tokens.begin(), r.tokens.begin() + keep);
    s.prompt_len = (int)keep;
    s.generated = 0;
    s.max_tokens = r.max_tokens;
    s.slot = b->free_slots.back();
    s.started = false;
    b->free_slots.pop_back();
    b->active.push_back(std::move(s));
}

extern "C" int b200_batch_create(b200_ctx* c, const b200_batch_config* cfg, b200_batch** out) {
    if (!c || !out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_create: null argument");
    if (!c->finalized) return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_create: context not finalized");
    if (c->par.world_size > 1) return fail(B200_ERR_UNSUPPORTED, "b200_batch_create: not available under tensor / expert parallelism");
    b200_batch* b = new b200_batch();
    b->c = c;
    b->cfg.max_batch_size = 8; b->cfg.max_seq_len = 4096; b->cfg.max_queue_depth = 64; b->cfg.eos_token_id = 2;   // BatchedEngineConfig::default (:32-40)
    if (cfg) b->cfg = *cfg;
    if (b->cfg.max_batch_size < 1 || b->cfg.max_batch_size > c->d.max_batch) {
        const int want = b->cfg.max_batch_size;
        delete b;
        return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_create: max_batch_size " + std::to_string(want) + " needs a context with as many sequence slots (max_batch = " +
                                                   std::to_string(c->d.max_batch) + ")");
    }
    if (b->cfg.max_seq_len < 2 || b->cfg.max_queue_depth < 1) {
        delete b;
        return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_create: bad max_seq_len / max_queue_depth");
    }
    for (int s = b->cfg.max_batch_size - 1; s >= 0; s--) b->free_slots.push_back(s);
    *out = b;
    return B200_OK;
}
extern "C" void b200_batch_destroy(b200_batch* b) { delete b; }

extern "C" int b200_batch_submit(b200_batch* b, const uint32_t* tokens, int n, int max_tokens, uint64_t* request_id) {
    if (!b || (n > 0 && !tokens) || n < 0 || max_tokens < 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_submit: bad argument");
    if (b->queue_count >= b->cfg.max_queue_depth) return fail(B200_ERR_OPERATION_FAILED, "queue full");
    for (int i = 0; i < n; i++)
        if (tokens[i] >= (uint32_t)b->c->d.vocab) return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_submit: token id exceeds vocab size");
    b->queue_count++;
    BatchReq r;
    r.id = b->next_id++;
    r.tokens.assign(tokens, tokens + n);
    r.max_tokens = max_tokens;
    if (request_id) *request_id = r.id;
    if ((int)b->active.size() < b->cfg.max_batch_size) batch_activate(b, std::move(r));
    else b->pending.push_back(std::move(r));
    return B200_OK;
}

static uint32_t host_argmax_last(const float* v, int n) {
    int bi = 0;
    float best = v[0];
    for (int i = 1; i < n; i++)
        if (v[i] >= best) { best = v[i]; bi = i; }
    return (uint32_t)bi;
}

extern "C" int b200_batch_step(b200_batch* b, b200_batch_event* out, int cap, int* n_out) {
    if (!b || (cap > 0 && !out) || cap < 0 || !n_out) return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_step: bad argument");
    b200_ctx* c = b->c;
    // step_sequence for every active sequence: who is finished, who prefills, who decodes
    const size_t n_act = b->active.size();
    std::vector<int> action(n_act, 0);   // 0 = finished before stepping, 1 = prompt, 2 = one token, 3 = error
    std::vector<uint32_t> next(n_act, 0);
    std::vector<std::string> err(n_act);
    std::vector<int> dec_idx, dec_slots;
    std::vector<uint32_t> dec_tokens;
    for (size_t i = 0; i < n_act; i++) {
        BatchSeq& s = b->active[i];
        if (!s.tokens.empty() && s.tokens.back() == b->cfg.eos_token_id) continue;   // :373-377
        if (s.generated >= s.max_tokens) continue;                                    // :379-381
        if (!s.started) { action[i] = 1; continue; }
        if (c->slots[s.slot].host_pos + 1 > (uint64_t)c->d.max_seq_len) { action[i] = 3; err[i] = "context length exceeded"; continue; }
        action[i] = 2;
        dec_idx.push_back((int)i);
        dec_slots.push_back(s.slot);
        dec_tokens.push_back(s.tokens.back());
    }
    // Prompts (ctx.position == 0 -> the whole prompt in this step, :383-389).  Long prompts of a dense model go through b200_prefill's
    // tensor-core pass one sequence at a time.  Short ones (below the GEMM prompt threshold) are fed as ROWS, one token per sequence
    // and sub-iteration, through the same batched pass that the decoding sequences use: 32 new 16-token prompts cost 16 passes over
    // the weights instead of 512 single-token launches.  The decoding sequences ride in sub-iteration 0.
    std::vector<int> short_idx, long_idx;
    size_t max_short = 0;
    for (size_t i = 0; i < n_act; i++) {
        if (action[i] != 1) continue;
        BatchSeq& s = b->active[i];
        const int rc = b200_reset(c, s.slot);
        if (rc) { action[i] = 3; err[i] = g_err; continue; }
        if ((int)s.tokens.size() >= c->prefill_gemm_min && prefill_gemm_ok(c, true)) long_idx.push_back((int)i);
        else { short_idx.push_back((int)i); max_short = std::max(max_short, s.tokens.size()); }
    }
    const size_t n_sub = std::max<size_t>(max_short, dec_idx.empty() ? 0 : 1);
    for (size_t k = 0; k < n_sub; k++) {
        std::vector<int> who, slots_k;
        std::vector<uint32_t> toks_k;
        if (k == 0)
            for (size_t q = 0; q < dec_idx.size(); q++) { who.push_back(dec_idx[q]); slots_k.push_back(dec_slots[q]); toks_k.push_back(dec_tokens[q]); }
        for (int i : short_idx) {
            BatchSeq& s = b->active[i];
            if (action[i] != 1 || k >= s.tokens.size()) continue;
            who.push_back(i); slots_k.push_back(s.slot); toks_k.push_back(s.tokens[k]);
        }
        if (who.empty()) continue;
        std::vector<uint32_t> picked(who.size());
        const int rc = b200_decode_batch_greedy(c, slots_k.data(), toks_k.data(), (int)who.size(), picked.data());
        for (size_t q = 0; q < who.size(); q++) {
            const int i = who[q];
            if (rc) { action[i] = 3; err[i] = g_err; continue; }
            if (action[i] == 2) next[i] = picked[q];
            else if (k + 1 == b->active[i].tokens.size()) { next[i] = picked[q]; b->active[i].started = true; }
        }
        b->batched_rows += who.size();
    }
    std::vector<float> logits;
    for (int i : long_idx) {
        BatchSeq& s = b->active[i];
        logits.resize((size_t)c->d.vocab);
        const int rc = b200_prefill(c, s.slot, s.tokens.data(), (int)s.tokens.size(), logits.data());
        if (rc) { action[i] = 3; err[i] = g_err; continue; }
        next[i] = host_argmax_last(logits.data(), c->d.vocab);
        s.started = true;
    }
    b->steps++;
    // results in the order of the active list, removals as Vec::remove (:241-289)
    std::vector<BatchSeq> keep;
    keep.reserve(n_act);
    for (size_t i = 0; i < n_act; i++) {
        BatchSeq& s = b->active[i];
        bool done = action[i] == 0;
        if (action[i] == 3) {
            batch_event(b, s.id, B200_BATCH_ERROR, 0, B200_FINISH_ERROR, s.prompt_len, s.generated);
            b->last_error = err[i];
            b->free_slots.push_back(s.slot);
            b->queue_count = std::max(0, b->queue_count - 1);
            continue;
        }
        if (action[i] == 1 || action[i] == 2) {
            s.tokens.push_back(next[i]);
            s.generated++;
            if (next[i] == b->cfg.eos_token_id) done = true;   // :397-399: EOS ends the sequence, no Token event for it
            else batch_event(b, s.id, B200_BATCH_TOKEN, next[i], 0, s.prompt_len, s.generated);
        }
        if (done) {
            batch_event(b, s.id, B200_BATCH_DONE, 0, s.generated >= s.max_tokens ? B200_FINISH_MAX_TOKENS : B200_FINISH_STOP, s.prompt_len, s.generated);
            b->free_slots.push_back(s.slot);
            b->queue_count = std::max(0, b->queue_count - 1);
            continue;
        }
        keep.push_back(std::move(s));
    }
    b->active.swap(keep);
    while ((int)b->active.size() < b->cfg.max_batch_size && !b->pending.empty()) {   // :291-304 (pending.pop(): newest first)
        BatchReq r = std::move(b->pending.back());
        b->pending.pop_back();
        batch_activate(b, std::move(r));
    }
    const int n = std::min<int>(cap, (int)b->events.size());
    for (int i = 0; i < n; i++) out[i] = b->events[i];
    b->events.erase(b->events.begin(), b->events.begin() + n);
    *n_out = n;
    return B200_OK;
}

extern "C" int b200_batch_counts(b200_batch* b, int* active, int* pending, int* undelivered_events, uint64_t* steps, uint64_t* decode_rows) {
    if (!b) return fail(B200_ERR_INVALID_ARGUMENT, "b200_batch_counts: null batch");
    if (active) *active = (int)b->active.size();
    if (pending) *pending = (int)b->pending.size();
    if (undelivered_events) *undelivered_events = (int)b->events.size();
    if (steps) *steps = b->steps;
    if (decode_rows) *decode_rows = b->batched_rows;
    return B200_OK;
}
extern "C" const char* b200_batch_last_error(b200_batch* b) { return b ? b->last_error.c_str() : ""; }


// ------------------------------------------------------------------ greedy continuation + lab positioning
extern "C" int b200_ctx_set_speculation(b200_ctx* c, int on) {
    if (!c) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_set_speculation: null ctx");
    if (!c->finalized) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_set_speculation: finalize the context first");
    for (int s = 0; s < (int)c->slots.size(); s++) {
        const int rc = spec_drain(c, s);
        if (rc) return rc;
    }
    if (on && !c->h_spec_pick) {
        CU(cudaSetDevice(c->par.device));
        for (int b = 0; b < 2; b++) {
            CU_ALLOC(cudaHostAlloc((void**)&c->h_spec_logits[b], (size_t)c->vocab_l * 4, cudaHostAllocDefault));
            CU(cudaEventCreateWithFlags(&c->spec_ev[b], cudaEventDisableTiming));
        }
        CU_ALLOC(cudaHostAlloc((void**)&c->h_spec_pick, 2 * sizeof(int), cudaHostAllocDefault));
    }
    c->speculate = on != 0;
    return B200_OK;
}
extern "C" int b200_ctx_speculation_stats(b200_ctx* c, int* enabled, uint64_t* hits, uint64_t* misses) {
    if (!c) return fail(B200_ERR_INVALID_ARGUMENT, "b200_ctx_speculation_stats: null ctx");
    if (enabled) *enabled = (c->speculate && c->mega_ok && c->par.world_size == 1 && !c->use_taps) ? 1 : 0;
    if (hits) *hits = c->spec_hits;
    if (misses) *misses = c->spec_misses;
    return B200_OK;
}
// Lab: declare the first `pos` positions of the slot's KV cache valid as they are (timing attention at depth without a prompt).
extern "C" int b200_debug_set_position(b200_ctx* c, int seq, uint64_t pos) {
    int rc;
    if ((rc = check_slot(c, seq, "b200_debug_set_position"))) return rc;
    if (pos >= (uint64_t)c->d.max_seq_len) return fail(B200_ERR_INVALID_ARGUMENT, "b200_debug_set_position: beyond the context");
    if ((rc = flush_pending(c, seq))) return rc;
    CU(cudaSetDevice(c->par.device));
    CU(cudaStreamSynchronize(c->stream));
    const int p = (int)pos;
    CU(cudaMemcpy(&c->slots[seq].d_state->pos_next, &p, sizeof(int), cudaMemcpyHostToDevice));
    c->slots[seq].host_pos = pos;
    return B200_OK;
}

// Host logic of the split-K planners, callable without a device (tests): K range per split (0 = unsplit) for an n_rows x K matrix at
// T token rows on n_sm SMs; persistent = the cost model of gemm_umma2.cuh's kernel, else round 1's rule.
extern "C" int b200_debug_plan_split(int n_rows, int K, int T, int n_sm, int persistent, int* k_split_out) {
    if (!k_split_out || n_rows <= 0 || K <= 0 || K % 256 || T <= 0 || n_sm <= 0) return fail(B200_ERR_INVALID_ARGUMENT, "b200_debug_plan_split: bad argument");
    UmmaParams p{};
    p.n_rows = n_rows; p.K = K; p.T = T;
    float dummy = 0.0f;
    if (persistent) umma2_plan_split(p, &dummy, (size_t)1 << 40, n_sm);
    else umma_plan_split(p, &dummy, (size_t)1 << 40, n_sm);
    *k_split_out = p.k_split;
    return B200_OK;
}
