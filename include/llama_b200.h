/* llama_b200.h — C ABI of the `cuda-b200` backend for Lexmata/llama-gguf.
 *
 * This is the drop-in boundary: a Rust shim binds these symbols
 * with `extern "C"` and implements the crate's `GpuInference` trait
 * (src/backend/mod.rs:283-296) and `Backend` trait (src/backend/mod.rs:29-265)
 * on top of them.  Plain pointers and sizes only; every host pointer is read or
 * written DURING the call and never retained (the Rust side owns `Vec<u8>`).
 *
 * All functions return 0 on success or a negative b200_status; the message is
 * available from b200_last_error() (thread-local).  No exceptions cross the ABI.
 */
#ifndef LLAMA_B200_H
#define LLAMA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* 1:1 with BackendError's variants (src/backend/error.rs:3-37). */
typedef enum b200_status {
    B200_OK = 0,
    B200_ERR_NOT_AVAILABLE = -1,         /* BackendError::NotAvailable */
    B200_ERR_SHAPE_MISMATCH = -2,        /* BackendError::ShapeMismatch */
    B200_ERR_DTYPE_MISMATCH = -3,        /* BackendError::DTypeMismatch */
    B200_ERR_UNSUPPORTED_DTYPE = -4,     /* BackendError::UnsupportedDType */
    B200_ERR_UNSUPPORTED = -5,           /* BackendError::Unsupported */
    B200_ERR_INVALID_ARGUMENT = -6,      /* BackendError::InvalidArgument */
    B200_ERR_TENSOR = -7,                /* BackendError::Tensor */
    B200_ERR_INITIALIZATION_FAILED = -8, /* BackendError::InitializationFailed */
    B200_ERR_ALLOCATION_FAILED = -9,     /* BackendError::AllocationFailed */
    B200_ERR_OPERATION_FAILED = -10      /* BackendError::OperationFailed */
} b200_status;

/* ggml type ids as stored in GGUF (src/gguf/constants.rs:56-89; DType mirrors
 * them, src/tensor/dtype.rs:175-210). */
enum {
    B200_TYPE_F32 = 0, B200_TYPE_F16 = 1, B200_TYPE_Q4_0 = 2, B200_TYPE_Q5_0 = 6, B200_TYPE_Q8_0 = 8,
    B200_TYPE_Q4_K = 12, B200_TYPE_Q5_K = 13, B200_TYPE_Q6_K = 14
};

/* What GpuOnlyInference::from_model reads out of ModelConfig
 * (src/backend/cuda/gpu_only.rs:426-520; src/model/loader.rs:62-300). */
typedef struct b200_model_desc {
    int32_t hidden;         /* {arch}.embedding_length */
    int32_t n_layers;       /* {arch}.block_count */
    int32_t n_heads;        /* {arch}.attention.head_count */
    int32_t n_kv_heads;     /* {arch}.attention.head_count_kv */
    int32_t head_dim;       /* key_length; 0 = hidden / n_heads (loader.rs:116-118) */
    int32_t ffn;            /* {arch}.feed_forward_length */
    int32_t vocab;          /* {arch}.vocab_size */
    int32_t max_seq_len;    /* min({arch}.context_length, EngineConfig::max_context_len) */
    float norm_eps;         /* attention.layer_norm_rms_epsilon (default 1e-5) */
    float rope_base;        /* rope.freq_base */
    float rope_scale;       /* rope.scale_linear; the position is DIVIDED by it (cpu/ops.rs:1300) */
    int32_t rope_neox;      /* 1 = pairs (i, i+hd/2) [qwen2]; 0 = pairs (2i, 2i+1) [llama] */
    int32_t n_experts;      /* 0 = dense FFN */
    int32_t n_experts_used; /* top-k */
    int32_t expert_ffn;     /* per-expert intermediate size (0 = ffn) */
    int32_t tied_output;    /* 1 = no output.weight; logits use token_embd (loader.rs:349-355) */
    int32_t max_batch;      /* number of independent sequence slots (>= 1) */
} b200_model_desc;

/* Tensor-parallel placement of this context: one context per rank/GPU.
 * Mirrors ShardingPlan::from_config's divisibility rules
 * (src/backend/tensor_parallel.rs:69-106). world_size == 1 -> no collectives. */
typedef struct b200_parallel_desc {
    int32_t world_size;
    int32_t rank;
    int32_t device; /* CUDA ordinal this rank drives */
} b200_parallel_desc;

typedef struct b200_ctx b200_ctx;

/* ---- library ------------------------------------------------------------ */
B200_API const char* b200_backend_name(void);       /* Backend::name() -> "cuda-b200" */
B200_API const char* b200_last_error(void);
B200_API int b200_device_count(int* out);           /* Backend::is_available() = (count > 0) */
B200_API int b200_type_block_elems(uint32_t ggml_type);
B200_API int b200_type_block_bytes(uint32_t ggml_type);

/* ---- GpuOnlyInference surface (src/backend/cuda/gpu_only.rs) ------------ */
/* from_model (gpu_only.rs:426): create -> upload every tensor -> finalize. */
B200_API int b200_ctx_create(const b200_model_desc* desc, const b200_parallel_desc* par /* NULL = 1 GPU, device 0 */,
                    b200_ctx** out);
/* One call per GGUF tensor, named as the loader names them (loader.rs:592-753,
 * 1140-1182): token_embd.weight, output_norm.weight, output.weight,
 * blk.N.{attn_norm,ffn_norm}.weight, blk.N.attn_{q,k,v,output}.weight,
 * blk.N.attn_{q,k,v}.bias, blk.N.ffn_{gate,up,down}.weight,
 * blk.N.ffn_gate_inp.weight, blk.N.ffn_{gate,up,down}_exps.weight.
 * ne[0] = in_features (contiguous), ne[1] = out_features, ne[2] = experts.
 * The FULL tensor is passed on every rank; the library keeps its own shard. */
B200_API int b200_ctx_upload_tensor(b200_ctx* ctx, const char* gguf_name, uint32_t ggml_type, const uint64_t* ne, int n_dims,
                           const void* host, size_t nbytes);
B200_API int b200_ctx_finalize(b200_ctx* ctx);
B200_API void b200_ctx_destroy(b200_ctx* ctx);

/* KV cache storage format (KVCacheFormat, src/model/kv_quantized.rs:11-20), chosen between b200_ctx_create and
 * b200_ctx_finalize (or with B200_KV_FORMAT=int8 in the environment): 0 = F32 (the KVCache of src/model/mod.rs:83-108, default),
 * 1 = Int8 -- QuantizedKVCache::write_kv / read_k_range semantics (kv_quantized.rs:143-216, 230-270, 366-392): one symmetric
 * scale = max|x| / 127 per (kv head, position) row, q = round(x / scale), attention over q * scale.  A quarter of the KV bytes; the
 * context then decodes on the per-op (graph) path and prefills token by token (B200_KV_INT8_GEMM=1: prompts of >= 32 tokens through
 * the tensor-core pass, outside the 1e-3 parity bound -- DESIGN 3.8).  The FP8 formats are not built (UNSUPPORTED). */
B200_API int b200_ctx_set_kv_format(b200_ctx* ctx, int format);
B200_API int b200_ctx_kv_format(b200_ctx* ctx, int* out);

/* GGUF -> HBM direct load.  Replaces GgufFile::open (src/gguf/mod.rs:23-40, mmap), GgufReader::read
 * (src/gguf/reader.rs:28-110: header, metadata, tensor infos, aligned data offset; versions 1-3), ModelLoader::parse_config
 * (src/model/loader.rs:62-300: the {arch}.* keys with the same defaults and the same strict typed getters, gguf/types.rs:78-97)
 * and the per-tensor Vec copy + upload of loader.rs:1341-1365 / cuda/gpu_only.rs:426-520: the file is mapped once, parsed in
 * place, and every tensor the engine uses goes from the page cache to its HBM allocation through two pinned staging buffers
 * (host threads fill one while the copy engine drains the other).  No tensor-sized host allocation, no f32 embedding table.
 * b200_gguf_* need no CUDA device.  Errors: bad magic / truncated file / missing key -> INVALID_ARGUMENT, version other than
 * 1-3 / architecture or feature outside the engine -> UNSUPPORTED, a used tensor of an unimplemented ggml type -> UNSUPPORTED_DTYPE. */
typedef struct b200_gguf b200_gguf;
typedef struct b200_load_stats {
    uint64_t file_bytes;      /* size of the mapping */
    uint64_t tensor_bytes;    /* bytes of the tensors handed to the engine (full tensors) */
    uint64_t device_bytes;    /* bytes this context copied to HBM (its shard under tensor / expert parallelism) */
    uint32_t tensors_loaded;
    uint32_t tensors_skipped; /* tensors of the file the engine has no use for (rope_freqs.weight ...) */
    double seconds;           /* wall time of the load loop including the final synchronisation of the load stream */
} b200_load_stats;
B200_API int b200_gguf_open(const char* path, b200_gguf** out);
B200_API void b200_gguf_close(b200_gguf* file);
B200_API int b200_gguf_info(b200_gguf* file, uint32_t* version, uint64_t* n_tensors, uint64_t* n_metadata, uint64_t* alignment,
                            uint64_t* data_offset, uint64_t* file_bytes);
B200_API int b200_gguf_architecture(b200_gguf* file, char* out, size_t cap);          /* general.architecture */
/* max_seq_len > 0: min(it, {arch}.context_length); 0: the file's context length (default 2048, loader.rs:131). */
B200_API int b200_gguf_model_desc(b200_gguf* file, int max_seq_len, int max_batch, b200_model_desc* out);
/* Tensor i: name (owned by `file`), ggml type, ne[4] (ne[0] contiguous), rank, pointer into the mapping, byte size.  *data is NULL
 * and *nbytes 0 for a ggml type the engine does not implement; every out pointer may be NULL. */
B200_API int b200_gguf_tensor_info(b200_gguf* file, uint64_t i, const char** name, uint32_t* ggml_type, uint64_t* ne4, int* n_dims,
                                   const void** data, size_t* nbytes);
/* Between b200_ctx_create and b200_ctx_finalize: upload every tensor of the file the engine has a slot for (this rank's shard). */
B200_API int b200_ctx_load_gguf(b200_ctx* ctx, b200_gguf* file, b200_load_stats* stats /* may be NULL */);
/* GpuOnlyInference::from_model for a path: open + describe + create + load + finalize.  kv_format as b200_ctx_set_kv_format. */
B200_API int b200_ctx_create_from_gguf(const char* path, const b200_parallel_desc* par /* NULL = 1 GPU, device 0 */, int max_seq_len,
                                       int max_batch, int kv_format, b200_ctx** out, b200_load_stats* stats /* may be NULL */);

/* Tensor parallel (replaces the gRPC all-reduce-via-rank-0 of
 * src/distributed/tensor_parallel_distributed.rs:135-187): one context per rank/GPU, created with
 * world_size/rank in b200_parallel_desc.  Between b200_ctx_create and b200_ctx_finalize every rank calls
 * b200_ctx_tp_handle (a 64-byte CUDA IPC handle of its exchange region), the host passes the handles around
 * (all-gather), and every rank calls b200_ctx_tp_set_peer for every other rank.  Uploads take the FULL tensor;
 * the library keeps this rank's column / row shard.  The kernels then exchange partial sums through peer memory
 * (NVLink).  b200_forward returns this rank's slice of the logits: [rank * vocab/P, (rank+1) * vocab/P). */
B200_API int b200_ctx_tp_handle(b200_ctx* ctx, void* handle_out64);
B200_API int b200_ctx_tp_set_peer(b200_ctx* ctx, int peer_rank, const void* handle64);

/* Single-process tensor / expert parallelism ("TP ranks are threads inside the library; callers never see ranks";
 * TensorParallel trait, src/backend/tensor_parallel.rs:13-32).  A group owns one context per device and one host thread per context;
 * peers are connected with cudaDeviceEnablePeerAccess (no IPC handles, no other process).  Dense models are tensor parallel
 * (ShardingPlan::from_config, tensor_parallel.rs:69-106), MoE models expert parallel.  Every call below fans out to the ranks and
 * returns when all of them are done; logits_out always receives the FULL `vocab` row.  devices == NULL: ordinals 0 .. n_devices-1.
 * n_devices == 1 is an ordinary single-GPU context behind the same interface.  The Rust shim creates a group when B200_TP is set:
 * nothing above GpuOnlyInference changes. */
typedef struct b200_group b200_group;
B200_API int b200_group_create(const b200_model_desc* desc, int n_devices, const int* devices, b200_group** out);
B200_API int b200_group_upload_tensor(b200_group* g, const char* gguf_name, uint32_t ggml_type, const uint64_t* ne, int n_dims,
                             const void* host, size_t nbytes);   /* the FULL tensor; every rank keeps its shard */
/* Every rank stages its own shard out of the shared mapping, in parallel (one host thread per device). */
B200_API int b200_group_load_gguf(b200_group* g, b200_gguf* file, b200_load_stats* stats /* rank 0's; may be NULL */);
B200_API int b200_group_finalize(b200_group* g);
B200_API void b200_group_destroy(b200_group* g);
B200_API int b200_group_forward(b200_group* g, int seq, uint32_t token, float* logits_out);
B200_API int b200_group_prefill_token(b200_group* g, int seq, uint32_t token);
B200_API int b200_group_reset(b200_group* g, int seq);
B200_API int b200_group_position(b200_group* g, int seq, uint64_t* out);
B200_API int b200_group_decode_greedy(b200_group* g, int seq, uint32_t first_token, int n_steps, uint32_t* tokens_out, float* elapsed_ms);
B200_API int b200_group_size(b200_group* g, int* out);
B200_API int b200_group_ctx(b200_group* g, int rank, b200_ctx** out);   /* rank's context (stats, path, debug); do not destroy */

/* GpuInference::forward (backend/mod.rs:285): one token through every layer;
 * logits_out receives `vocab` f32 on the host. */
B200_API int b200_forward(b200_ctx* ctx, int seq, uint32_t token, float* logits_out);
/* GpuInference::prefill_token (backend/mod.rs:289): KV update, no logits. */
B200_API int b200_prefill_token(b200_ctx* ctx, int seq, uint32_t token);
/* GpuOnlyInference::forward_batch (gpu_only.rs:776-790): n tokens of ONE
 * sequence; logits of the last one (logits_out may be NULL). */
B200_API int b200_prefill(b200_ctx* ctx, int seq, const uint32_t* tokens, int n, float* logits_out);
/* One token for each of n DISTINCT sequence slots;
 * logits_out is n x vocab. */
B200_API int b200_decode_batch(b200_ctx* ctx, const int* seqs, const uint32_t* tokens, int n, float* logits_out);
/* Greedy continuation behind b200_forward (opt-in, after finalize; one GPU, megakernel paths).  When on, b200_forward also picks
 * argmax (last maximum wins, src/main.rs:1816-1821) on the device and launches the NEXT token on that pick before it returns.  If
 * the next b200_forward of the sequence passes exactly that token -- what every greedy caller does, e.g. the reference's bench
 * loop and a Sampler at temperature 0 -- its logits are already on their way and the host's turnaround (D2H, argmax, call
 * overhead) overlaps the next token instead of idling the GPU; any other token, or any other entry point on the sequence, waits
 * for the token in flight, discards it (the device position goes back; its KV rows lie beyond the position) and proceeds as
 * usual.  Results are identical either way.  b200_position never counts a token in flight. */
B200_API int b200_ctx_set_speculation(b200_ctx* ctx, int on);
B200_API int b200_ctx_speculation_stats(b200_ctx* ctx, int* enabled, uint64_t* hits, uint64_t* misses);

/* b200_decode_batch with the greedy pick (last maximum wins, src/main.rs:1816-1821) made on the device: next_out receives n token
 * ids instead of n x vocab logits. */
B200_API int b200_decode_batch_greedy(b200_ctx* ctx, const int* seqs, const uint32_t* tokens, int n, uint32_t* next_out);

/* Continuous batching: the peer of BatchedEngine (src/engine_batched.rs) below the tokenizer and the channels.
 * Requests become sequences that each own one KV slot of the context; b200_batch_step is one iteration of the reference's
 * background loop (engine_batched.rs:199-320) with step_sequence's rules (:366-411) -- the whole prompt on a sequence's first
 * step, one token per step after that, finished when the last token is EOS or max_tokens were generated, EOS itself is not
 * reported as a token, pending requests are promoted newest-first (Vec::pop, :291-304) -- except that all decoding sequences of
 * a step share ONE pass over the weights (b200_decode_batch_greedy) instead of one Model::forward each.  Greedy sampling.
 * The config's defaults are BatchedEngineConfig::default (:32-40); max_batch_size <= the context's max_batch. */
typedef struct b200_batch b200_batch;
typedef struct b200_batch_config {
    int32_t max_batch_size;   /* concurrent sequences (8) */
    int32_t max_seq_len;      /* prompts are cut to max_seq_len - 1 tokens (4096; never above the context's) */
    int32_t max_queue_depth;  /* active + pending requests beyond which submit fails with "queue full" (64) */
    uint32_t eos_token_id;
} b200_batch_config;
enum { B200_BATCH_TOKEN = 0, B200_BATCH_DONE = 1, B200_BATCH_ERROR = 2 };              /* BatchToken (:60-72) */
enum { B200_FINISH_STOP = 0, B200_FINISH_MAX_TOKENS = 1, B200_FINISH_ERROR = 2 };      /* BatchFinishReason (:75-79) */
typedef struct b200_batch_event {
    uint64_t request_id;
    int32_t kind;               /* B200_BATCH_* */
    uint32_t token;             /* TOKEN: the generated id */
    int32_t reason;             /* DONE: B200_FINISH_* */
    int32_t prompt_tokens;      /* DONE: as BatchToken::Done */
    int32_t completion_tokens;
} b200_batch_event;
B200_API int b200_batch_create(b200_ctx* ctx, const b200_batch_config* cfg /* NULL = defaults */, b200_batch** out);
B200_API void b200_batch_destroy(b200_batch* batch);
/* BatchedEngine::submit (:167-192): OPERATION_FAILED "queue full" past max_queue_depth; an empty prompt is accepted and answered
 * with an ERROR event ("empty prompt", :329-334). */
B200_API int b200_batch_submit(b200_batch* batch, const uint32_t* tokens, int n, int max_tokens, uint64_t* request_id);
/* One loop iteration; up to `cap` events (oldest first) are copied out, the rest wait for the next call. */
B200_API int b200_batch_step(b200_batch* batch, b200_batch_event* events, int cap, int* n_events);
B200_API int b200_batch_counts(b200_batch* batch, int* active, int* pending, int* undelivered_events, uint64_t* steps,
                               uint64_t* decode_rows);
B200_API const char* b200_batch_last_error(b200_batch* batch);   /* message of the most recent ERROR event */

/* GpuInference::reset / position (backend/mod.rs:292-295). */
B200_API int b200_reset(b200_ctx* ctx, int seq);
B200_API int b200_position(b200_ctx* ctx, int seq, uint64_t* out);

/* Device-resident greedy decode (sampling on device).
 * Runs n_steps tokens starting from `first_token`, argmax (last max wins,
 * src/main.rs:1816-1821) on the GPU, no host round trip between tokens.
 * tokens_out (n_steps, may be NULL) receives the generated ids; elapsed_ms
 * (may be NULL) the CUDA-event time of the n_steps on the launching stream. */
B200_API int b200_decode_greedy(b200_ctx* ctx, int seq, uint32_t first_token, int n_steps, uint32_t* tokens_out,
                       float* elapsed_ms);

/* Debug/parity taps: hidden state of the last processed token after `layer`
 * layers (0 = embedding, n_layers = input of the final norm). */
B200_API int b200_get_hidden(b200_ctx* ctx, int seq, int layer, float* out);
/* Debug: per-phase globaltimer stamps of the per-token megakernel (first call arms them; see csrc/mega.cuh). */
B200_API int b200_debug_mega_timeline(b200_ctx* ctx, unsigned long long* out, int max_n);
B200_API int b200_debug_read(b200_ctx* ctx, int which, float* out, int n);
B200_API int b200_debug_mega_phase(b200_ctx* ctx, int phase, unsigned long long* out, int max_n);
/* Watchdog words of the tensor-pipe / megakernel paths: out8[0] = 0 when no bounded wait ever gave up, else
 * (code, which wait, CTA, sequence number); clears them. */
B200_API int b200_debug_err(b200_ctx* ctx, int* out8);
/* Host logic of the dequant-GEMM's split-K planners (no device needed): K range per split, 0 = unsplit. */
B200_API int b200_debug_plan_split(int n_rows, int K, int T, int n_sm, int persistent, int* k_split_out);
/* Lab: take the first `pos` positions of the slot's KV cache as valid as they are (attention at depth without a prompt). */
B200_API int b200_debug_set_position(b200_ctx* ctx, int seq, uint64_t pos);
/* Decode path chosen by b200_ctx_finalize: 0 = CUDA graph of per-op kernels, 1 = per-token megakernel,
 * 2 = streamed megakernel (TMA producer warp + mbarrier ring, csrc/stream.cuh). */
B200_API int b200_ctx_path(b200_ctx* ctx, int* out);
/* Statistics for bench.py: kernels launched by this library since creation. */
B200_API int b200_ctx_stats(b200_ctx* ctx, uint64_t* kernel_launches, uint64_t* weight_bytes, uint64_t* kv_bytes_per_pos);
/* Roofline probe: replays only the dequant-GEMV launches of one token (same arguments and order
 * as the decode graph) `iters` times between CUDA events; bytes = weight bytes those launches read. */
B200_API int b200_bench_gemv_pass(b200_ctx* ctx, int seq, int iters, float* avg_ms_per_pass, uint64_t* launches_per_pass,
                         uint64_t* bytes_per_pass);
/* Times `iters` launches of the vec_mat_q kernel over an uploaded weight with
 * CUDA events on the launching stream (per-kernel GB/s for bench.py). */
B200_API int b200_bench_weight_gemv(b200_ctx* ctx, const char* gguf_name, int iters, float* avg_ms, uint64_t* bytes);

/* ---- Backend per-op surface (src/backend/mod.rs:29-265) ------------------
 * Host pointers in, host pointers out: H2D, kernel, D2H, like the existing
 * CUDA per-op path (src/backend/cuda/mod.rs:229-835).  f32 activations. */
B200_API int b200_op_add(const float* a, const float* b, float* out, size_t n);
B200_API int b200_op_mul(const float* a, const float* b, float* out, size_t n);
B200_API int b200_op_scale(const float* a, float s, float* out, size_t n);
B200_API int b200_op_silu(const float* x, float* out, size_t n);
B200_API int b200_op_gelu(const float* x, float* out, size_t n);
B200_API int b200_op_softmax(const float* x, float* out, size_t n);
B200_API int b200_op_rms_norm(const float* x, const float* weight, float eps, float* out, size_t n_rows, size_t hidden);
/* matmul: a[m][k] @ b[k][n] -> out[m][n], row-major f32 (Backend::matmul, src/backend/mod.rs:90; cpu/ops.rs:429-487).
 * matvec: a[m][k] @ b[k] -> out[m] (mod.rs:93; cpu/ops.rs:531-575).  matvec_q: a = m rows of k/bs quantised blocks (mod.rs:107;
 * cpu/ops.rs:922-946).  Required trait methods; the model code never calls them (compatibility surface). */
B200_API int b200_op_matmul(const float* a, const float* b, float* out, size_t m, size_t k, size_t n);
B200_API int b200_op_matvec(const float* a, const float* b, float* out, size_t m, size_t k);
B200_API int b200_op_matvec_q(const void* a, uint32_t ggml_type, const float* b, float* out, size_t m, size_t k);
/* vec_mat: a[k] @ W[k,n] (f32 W, GGUF layout: n rows of k). */
B200_API int b200_op_vec_mat(const float* a, const float* w, float* out, size_t k, size_t n);
/* vec_mat_q: W quantised, n rows of k/bs blocks (backend/mod.rs:110). */
B200_API int b200_op_vec_mat_q(const float* a, const void* w, uint32_t ggml_type, float* out, size_t k, size_t n);
/* Extension (no counterpart in the reference's Backend trait, which runs a prompt as T vec_mat_q calls,
 * src/model/llama.rs:327-345): T rows at once through the tcgen05/TMEM dequant-GEMM.  a [t_rows][k] f32,
 * W = n rows of k/bs GGUF blocks (Q4_K, Q5_K, Q6_K, Q8_0), out [t_rows][n].  fp16 operands, f32 accumulation. */
B200_API int b200_op_mat_mat_q(const float* a, const void* w, uint32_t ggml_type, float* out, size_t t_rows, size_t k, size_t n);
/* dequantize: bit-exact with the reference's dequantize_* (tensor/quant/dequant.rs). */
B200_API int b200_op_dequantize(const void* src, uint32_t ggml_type, float* out, size_t n_elems);
/* rope: q[n_heads,1,hd], k[n_kv,1,hd] in place (cpu/ops.rs:1216-1337). */
B200_API int b200_op_rope(float* q, float* k, int n_heads, int n_kv_heads, int head_dim, int pos, float freq_base,
                 float freq_scale, int use_neox);
/* attention_cached: q[n_heads,1,hd]; k/v cache [n_kv,max_seq,hd] (cpu/ops.rs:1479-1537). */
B200_API int b200_op_attention_cached(const float* q, const float* k_cache, const float* v_cache, float* out, int n_heads,
                             int n_kv_heads, int head_dim, int max_seq, float scale, int kv_len);
/* attention: causal, q[n_heads,seq,hd], k/v[n_kv,seq,hd] (backend/mod.rs attention). */
B200_API int b200_op_attention(const float* q, const float* k, const float* v, float* out, int n_heads, int n_kv_heads,
                      int seq_len, int head_dim, float scale);

#ifdef __cplusplus
}
#endif
#endif /* LLAMA_B200_H */
