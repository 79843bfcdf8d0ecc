// mega.cuh — one persistent, cooperative kernel per decoded token (or per run of greedy tokens).
//
// The graph path launches ~195 kernels per Llama-3-8B token; in-kernel timelines (profiles/r01_v4_gemv_mma.md)
// show ~5 us of fixed cost per launch (launch, x staging behind cold loads, merge, exit) against 1.5-10 us of
// streaming.  Here the token is ONE launch of one CTA per SM; the layer sequence of GpuOnlyInference::forward
// (src/backend/cuda/gpu_only.rs:849-1010; CPU: model/llama.rs:275-362, layers.rs:1082-1245) becomes a list of
// phases separated by grid barriers:
//
//   embed -> L x { QKV GEMV | RoPE + KV write + GQA attention | O GEMV + residual |
//                  gate/up GEMV + SwiGLU | down GEMV + residual } -> vocab-head GEMV [-> argmax -> next token]
//
// GEMV phases are mma_gemv_cta (gemv_mma.cuh) with its arguments read from a per-phase descriptor in global
// memory; before a CTA waits at a barrier it pulls the first units of the NEXT GEMV towards L2, so the weight
// stream does not stop at a phase boundary.  RoPE and the KV write are folded into the attention phase
// (attention.cuh: attn_decode_item with qkv_raw), which removes two launches per layer.
//
// Dense models whose every weight launch is eligible for the tensor-pipe GEMV take this path (engine.cu:
// mega_build); MoE models and the 32-element block types stay on the CUDA-graph path.
#pragma once
#include "attention.cuh"
#include "gemv_mma.cuh"
#include "misc.cuh"

namespace b200 {

enum : int { PH_GEMV = 0, PH_ATTN = 1 };
enum : int { MEGA_LOGITS = 0, MEGA_PREFILL = 1, MEGA_GREEDY = 2 };

struct alignas(16) MegaPhase {   // 16-byte multiple: stream2.cuh moves descriptors with cp.async.bulk
    int kind;
    int tp_sync;   // tensor parallel: the phase wrote partial sums to the peers; exchange flags before the next one
    int pad[2];
    MParams gemv;
    AttnParams attn;
};

struct MegaParams {
    const MegaPhase* phases;  // device array: 5 per layer, then the vocab head (last)
    int n_phases;
    int mode;
    int n_tokens;             // tokens processed by this launch (MEGA_GREEDY: argmax feeds the next one)
    unsigned int* bar;        // grid-barrier counter, zeroed by the host before every launch
    int* err;                 // watchdog flag (2 = a grid barrier timed out)
    // embedding row -> residual stream
    int embd_type;
    const uint8_t* embd;
    long long embd_row_bytes;
    int hidden, vocab;
    float* h;
    // sampling state
    SeqState* st;
    const float* logits;
    float* cand_val;          // [grid] per-CTA argmax candidates
    int* cand_idx;
    int* generated;
    int max_generated;
    int hd, G;
    int early;                // issue a GEMV phase's first weight copies before waiting at the barrier that precedes it
    unsigned long long* dbg;  // optional [n_phases + 3] globaltimer stamps of CTA 0 for the LAST token of the launch
    // tensor parallel: one megakernel per rank/GPU, partial sums and flags travel through peer memory (NVLink)
    int tp_size, tp_rank;
    int vocab_local;                          // rows of the vocab head owned by this rank (argmax offset = rank * vocab_local)
    unsigned int* tp_flags;                   // local [tp_size]: slot r is written by rank r (monotonic epochs)
    unsigned int* tp_peer_flags[kMmaMaxPeers];  // the same array on every rank, as mapped here
    unsigned int tp_epoch0;                   // exchanges of this launch use epochs tp_epoch0 + 1, + 2, ...
    float* tp_cand;                           // local [tp_size][2]: (best logit, global index as float bits) of every rank
    float* tp_peer_cand[kMmaMaxPeers];
};

// Cross-GPU flag exchange after the local grid barrier: CTA 0 tells every peer "my partial sums of exchange `epoch`
// are in your memory"; every CTA then waits until all ranks have said so.  Bounded spin.
__device__ __forceinline__ bool tp_exchange(const MegaParams& mp, unsigned int epoch, int* s_flag) {
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid < mp.tp_size)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mp.tp_peer_flags[tid] + mp.tp_rank), "r"(epoch) : "memory");
    if (tid == 0) {
        int ok = 1;
        const long long t0 = clock64();
        for (int r = 0; r < mp.tp_size && ok; r++) {
            for (;;) {
                unsigned int v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mp.tp_flags + r) : "memory");
                if ((int)(v - epoch) >= 0) break;
                if (clock64() - t0 > 3000000000LL) {
                    ok = 0;
                    atomicExch(mp.err, 3);
                    break;
                }
            }
        }
        *s_flag = ok;
    }
    __syncthreads();
    return *s_flag != 0;
}

// Grid barrier on a monotonically increasing counter, split in two so that work that does not depend on other
// CTAs (fetching the descriptor of a later phase) runs while thread 0 waits.  Release/acquire at gpu scope:
// everything the CTA wrote before grid_arrive is visible to every CTA after grid_wait.  Bounded spin: a lost CTA
// must never hang the box.
__device__ __forceinline__ void grid_arrive(unsigned int* bar, unsigned int& target) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
// grid_wait_busy: the warps that do not poll call busy() (one small piece of independent work per call, e.g. an L2
// prefetch of the next weights) until it returns false or thread 0 has seen the barrier open (s_open == target).
template <class Busy>
__device__ __forceinline__ bool grid_wait_busy(unsigned int* bar, unsigned int target, int* err, int* s_flag, volatile unsigned int* s_open,
                                               Busy&& busy) {
    if (threadIdx.x == 0) {
        unsigned int v = 0;
        int ok = 1;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 3000000000LL) {  // ~1.5 s
                ok = 0;
                atomicExch(err, 2);
                break;
            }
        }
        *s_flag = ok;
        *s_open = target;
    } else if (threadIdx.x >= 32) {
        while (*s_open != target && busy()) {}
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ bool grid_wait(unsigned int* bar, unsigned int target, int* err, int* s_flag) {
    if (threadIdx.x == 0) {
        unsigned int v = 0;
        int ok = 1;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 3000000000LL) {  // ~1.5 s
                ok = 0;
                atomicExch(err, 2);
                break;
            }
        }
        *s_flag = ok;
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ bool grid_barrier(unsigned int* bar, unsigned int& target, int* err, int* s_flag) {
    grid_arrive(bar, target);
    return grid_wait(bar, target, err, s_flag);
}

// One instantiation per attention shape (head_dim, query heads per kv head rounded up to 4 / 8): a model only ever
// runs one of them, and the kernel's code footprint matters (every phase starts on a cold instruction path).
template <int HD, int GMAX>
__global__ void __launch_bounds__(kMmaMaxWarps * 32, 1) mega_decode_kernel(const __grid_constant__ MegaParams mp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float s_red[2 * kMmaMaxWarps];
    __shared__ float s_part[kMmaMaxWarps][2][2][32];
    __shared__ unsigned int s_ticket;
    __shared__ int s_flag;
    __shared__ unsigned int s_open;
    __shared__ __align__(16) MegaPhase s_phs[3];   // descriptors of the running phase and the next two (ring)
    __shared__ float s_rope[HD];   // per-token RoPE table: cos(theta_i) at [i], sin(theta_i) at [HD/2 + i]
    __shared__ float s_av[kMmaMaxWarps];
    __shared__ int s_ai[kMmaMaxWarps];

    constexpr int NW = kMmaMaxWarps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int target = 0;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;  // prefill: no vocab head

    // descriptor ring: slot (global phase number % 3).  Threads of warps >= 1 fetch descriptor `ph` (of the token's
    // program, cyclic) while thread 0 spins in a barrier; it is used two barriers later.
    auto fetch_desc = [&](long long gph) {
        if (tid < 32) return;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(mp.phases + (int)(gph % n_run));
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_phs[gph % 3]);
        for (int i = tid - 32; i < (int)(sizeof(MegaPhase) / 4); i += (NW - 1) * 32) dst[i] = src[i];
    };
    fetch_desc(0);
    fetch_desc(1);
    if (tid == 0) s_open = 0u;
    __syncthreads();  // phase 0 looks at its descriptor before it reaches a barrier
    long long gph = 0;  // phases executed so far in this launch
    unsigned int tp_n = 0;  // cross-GPU exchanges so far in this launch

    for (int tok = 0; tok < mp.n_tokens; tok++) {
        // ---- embedding (LlamaModel::forward, model/llama.rs:293-306): CTA 0 dequantises row `token`, bit-exactly ----
        if (blockIdx.x == 0) {
            int token = mp.st->token;
            token = min(max(token, 0), mp.vocab - 1);
            const int be = type_block_elems(mp.embd_type), bb = type_block_bytes(mp.embd_type);
            const uint8_t* row = mp.embd + (long long)token * mp.embd_row_bytes;
            for (int i = tid; i < mp.hidden; i += NW * 32) {
                const int blk = i / be;
                mp.h[i] = dequant_elem(mp.embd_type, row + (long long)blk * bb, i - blk * be);
            }
            if (tid == 0) {
                const int pn = mp.st->pos_next;
                mp.st->pos_cur = pn;
                mp.st->pos_next = pn + 1;
            }
        }
        // Every phase BEGINS with the grid barrier that ends its predecessor (the first one: the embedding).  A GEMV
        // phase arrives, then deals its units and issues its first weight copies, and only then waits: the copies
        // and the descriptor fetch overlap the barrier latency.
        bool ok = true;
        bool prev_tp_sync = false;   // the predecessor left partial sums in peer memory
        int stamp = 0;
        auto bar_arrive = [&]() {
            if (prev_tp_sync) {  // the CTA's stores to peer memory must be visible system-wide before the barrier says so
                __syncthreads();
                if (tid == 0) asm volatile("fence.acq_rel.sys;" ::: "memory");
            }
            grid_arrive(mp.bar, target);
        };
        auto bar_wait = [&](auto&& busy) {
            fetch_desc(gph + 2);
            ok = grid_wait_busy(mp.bar, target, mp.err, &s_flag, &s_open, busy);
            if (ok && prev_tp_sync) ok = tp_exchange(mp, mp.tp_epoch0 + (++tp_n), &s_flag);
            if (mp.dbg && blockIdx.x == 0 && tid == 0) mp.dbg[stamp] = gtimer();
            stamp++;
        };

        for (int ph = 0; ph < n_run; ph++, gph++) {
            const MegaPhase& cur = s_phs[gph % 3];
            if (cur.kind == PH_GEMV) {
                // The GEMV that feeds an attention phase (QKV): while its barrier is closed, the waiting warps first pull
                // the K/V rows that attention will read (positions 0..pos of this layer, cold in HBM: the weight stream
                // has flushed L2 since the previous token) towards L2, then the GEMV's own next weights.
                const MegaPhase& nxt = s_phs[(gph + 1) % 3];
                #ifndef B200_KV_PF
#define B200_KV_PF 0
#endif
                const bool kv_pf = B200_KV_PF && (ph + 1 < n_run) && nxt.kind == PH_ATTN && tid >= 32;
                // (real loads, results discarded: a prefetch instruction is dropped on a TLB miss, and the page walk is most
                // of the latency of the first K/V access of a layer).  Position-major caches: rows 0..pos are contiguous.
                long long kv_li = (long long)blockIdx.x * ((NW - 1) * 32) + (tid - 32);
                const long long kv_lines = (kv_pf && nxt.attn.kv_pos_stride) ? ((long long)(*nxt.attn.pos + 1) * nxt.attn.kv_pos_stride) / 32 : 0;
                auto wait_kv = [&](auto&& busy) {
                    bar_wait([&]() -> bool {
                        if (kv_li < kv_lines) {
                            unsigned int sink;
                            asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(sink) : "l"(nxt.attn.k_cache + 32 * kv_li));
                            kv_li += (long long)gridDim.x * ((NW - 1) * 32);
                            return true;
                        }
                        return busy();
                    });
                };
                mma_gemv_cta(cur.gemv, smem, s_red, s_part, bar_arrive, wait_kv, mp.early != 0, false);
                if (!ok) return;
            } else {
                attn_stamp(cur.attn, 0);
                bar_arrive();
                bar_wait([] { return false; });
                if (!ok) return;
                const AttnParams& ap = cur.attn;
                attn_stamp(ap, 1);
                const int kv_len = *ap.pos + 1;
                if (ph == 1) {   // first attention phase of the token: the rotation angles (Backend::rope, cpu/ops.rs:1216-1337)
                    const float position = (float)(kv_len - 1) / ap.rope_scale;
                    for (int pi = tid; pi < HD / 2; pi += NW * 32) {
                        const float theta = position * ap.freq[pi];
                        s_rope[pi] = cosf(theta);
                        s_rope[HD / 2 + pi] = sinf(theta);
                    }
                    __syncthreads();
                }
                const int n_items = ap.n_kv * ap.n_splits;
                for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                    const int kh = item / ap.n_splits, split = item - kh * ap.n_splits;
                    float* sm = reinterpret_cast<float*>(smem);
                    attn_decode_item<HD, GMAX, NW>(ap, kh, split, kv_len, sm, &s_ticket, s_rope);
                    __syncthreads();
#ifdef B200_ATTN_TWICE   // experiment: the same item again with a warm instruction cache (single-split configs only)
                    attn_stamp(ap, 4);
                    attn_decode_item<HD, GMAX, NW>(ap, kh, split, kv_len, sm, &s_ticket, s_rope);
                    __syncthreads();
                    attn_stamp(ap, 5);
#endif
                }
            }
            prev_tp_sync = cur.tp_sync != 0;
        }
        // the barrier that ends the last phase of the token
        gph--;  // fetch_desc(gph + 2) inside bar_wait: keep the ring consistent with the loop above
        bar_arrive();
        bar_wait([] { return false; });
        gph++;
        if (!ok) return;

        if (mp.mode != MEGA_GREEDY) continue;
        // ---- greedy pick on the device: raw-logit argmax, LAST maximal index wins (src/main.rs:1816-1821) ----
        {
            const int per = (mp.vocab_local + gridDim.x - 1) / gridDim.x;
            const int lo = blockIdx.x * per, hi = min(mp.vocab_local, lo + per);
            float best = -INFINITY;
            int bi = -1;
            for (int i = lo + tid; i < hi; i += NW * 32) {
                const float v = mp.logits[i];
                if (v >= best || bi < 0) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_av[warp] = best; s_ai[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < NW; w++)
                    if (s_ai[w] >= 0 && (bi < 0 || s_av[w] > best || (s_av[w] == best && s_ai[w] > bi))) { best = s_av[w]; bi = s_ai[w]; }
                mp.cand_val[blockIdx.x] = best;
                mp.cand_idx[blockIdx.x] = bi;
            }
        }
        if (!grid_barrier(mp.bar, target, mp.err, &s_flag)) return;
        const unsigned int cand_epoch = mp.tp_epoch0 + (mp.tp_size > 1 ? ++tp_n : 0u);  // uniform over the grid
        if (blockIdx.x == 0) {
            if (warp == 0) {
                float best = -INFINITY;
                int bi = -1;
                for (int c = lane; c < (int)gridDim.x; c += 32) {
                    const float v = mp.cand_val[c];
                    const int i = mp.cand_idx[c];
                    if (i >= 0 && (bi < 0 || v > best || (v == best && i > bi))) { best = v; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
                }
                if (mp.tp_size > 1) {  // every rank picks the same winner among the per-rank candidates (ties: largest index)
                    bi += mp.tp_rank * mp.vocab_local;
                    if (lane < mp.tp_size) {
                        float* dst = mp.tp_peer_cand[lane] + 2 * mp.tp_rank;
                        asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(best) : "memory");
                        asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + 1), "f"(__int_as_float(bi)) : "memory");
                    }
                    __syncwarp();
                    asm volatile("fence.acq_rel.sys;" ::: "memory");
                    const unsigned int epoch = cand_epoch;
                    if (lane < mp.tp_size)
                        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mp.tp_peer_flags[lane] + mp.tp_rank), "r"(epoch) : "memory");
                    const long long t0 = clock64();
                    bool ok = true;
                    if (lane < mp.tp_size) {
                        for (;;) {
                            unsigned int v;
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mp.tp_flags + lane) : "memory");
                            if ((int)(v - epoch) >= 0) break;
                            if (clock64() - t0 > 3000000000LL) { ok = false; atomicExch(mp.err, 3); break; }
                        }
                    }
                    __syncwarp();
                    best = -INFINITY;
                    bi = -1;
                    if (ok && lane < mp.tp_size) {
                        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(best) : "l"(mp.tp_cand + 2 * lane) : "memory");
                        float fi;
                        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(fi) : "l"(mp.tp_cand + 2 * lane + 1) : "memory");
                        bi = __float_as_int(fi);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
                    }
                }
                if (lane == 0) {
                    mp.st->token = bi;
                    const int gcount = mp.st->n_generated;
                    if (gcount < mp.max_generated) mp.generated[gcount] = bi;
                    mp.st->n_generated = gcount + 1;
                }
            }
            __syncthreads();  // CTA 0 embeds the new token at the top of the loop: it must see st->token
        }
    }
}

}  // namespace b200
