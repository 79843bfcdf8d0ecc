// stream2.cuh — second streamed per-token megakernel: the round-2 decode path.
//
// Round 1 (stream.cuh) moved exactly the algorithmic bytes but left HBM idle two thirds of the time
// (profiles/r01_stream_ncu_full.md): ~4 us of fixed cost per phase (CTA-wide barriers around the grid
// barrier, the staged-x copy, merge + epilogue), whole 32-row tiles per CTA (448 gate/up tiles on 148 CTAs: the four heavy
// CTAs finish 4.4 us late; 128 O / down tiles leave 20 CTAs idle), 7 consumer warps at 255 registers (1.75 warps per
// scheduler) and 15.6 us of attention per layer behind six dependent global round trips.  This kernel keeps the idea
// (a TMA producer that streams weights across phase boundaries into an mbarrier ring) and changes everything around it:
//
//   * 16 warps at <= 128 registers: 14 CONSUMER warps (units2.cuh: one K half of B operands live at a time), one LOADER
//     warp, one PRODUCER warp.  No CTA-wide barrier anywhere in a GEMV phase.
//   * unit-balanced stream-K: the E = tiles x entries-per-tile ring entries of a phase are dealt to CTAs as contiguous,
//     equal ranges [b E / n, (b + 1) E / n), entry i of a CTA's range to warp PAIR i mod 7 (= ring order); the even warp
//     of a pair computes the first 128 elements of every 256-element chunk of the entry, the odd warp the second (the
//     unit kernels are organised by K halves, units2.cuh), so work is dealt in half-entries.  A 32-row
//     tile cut by a CTA boundary is finished by the CTA that holds its HEAD (which it reaches LAST); the CTAs holding the
//     tail pieces reach them FIRST and publish 32 partial sums as 8-byte (value, epoch) packets that the owner polls
//     (no flag, no fence: one round trip, long since landed when the owner gets there).
//   * inside a CTA a tile's pieces meet in shared memory: every contributing warp stores its 32 row sums and bumps a
//     counter; the LAST warp to arrive adds them in entry order (fixed: results are run-to-run identical), runs the
//     epilogue (bias, residual, SwiGLU, staged int8 form of the output) and moves on.  Four tile slots, generation-checked.
//   * the loader warp owns the phase boundary: it waits for the CTA's consumers on an mbarrier, arrives at the grid
//     barrier (one red.release.gpu per CTA), polls it, and pulls the phase's input into shared memory with ONE bulk copy
//     (cp.async.bulk, SASS UBLKCP) that completes on the mbarrier the consumers wait on; the next phase's descriptor
//     rides on the same mbarrier.  Consumers go from their last entry of phase p straight to the set-up of phase p + 1.
//   * attention (attn2_phase): K/V rows of earlier positions are requested BEFORE the boundary is waited for (they do
//     not depend on this token), q / k / v of the new position arrive through the loader's bulk copies and are rotated
//     in shared memory; the new position's row is consumed from shared memory while it is written to the cache; short
//     contexts use one CTA per kv head (no cross-CTA merge at all), long ones split and merge through a ticket.
//   * the embedding row, its staged form and (greedy) the pick of the next token are a phase of their own (CTA 0); argmax
//     candidates are collected in the vocab head's epilogue (no extra pass over the logits, no extra grid barrier).
//
// Replaces the same reference code as stream.cuh: GpuOnlyInference::forward (src/backend/cuda/gpu_only.rs:849-1010), CPU
// LlamaModel::forward (src/model/llama.rs:275-362) with the fused dots of src/backend/cpu/simd.rs:931-1146, RoPE
// (cpu/ops.rs:1216-1337), KV write (model/layers.rs:580-600), attention_cached (cpu/ops.rs:1479-1537), greedy rule
// (src/main.rs:1816-1821).
#pragma once
#include "stream.cuh"
#include "units2.cuh"

namespace b200 {

constexpr int kS2Cons = 14;                       // consumer warps
constexpr int kS2Pairs = kS2Cons / 2;              // a ring entry is computed by a PAIR of warps: even warp K half 0, odd warp K half 1
constexpr int kS2NT = kS2Cons * 32;               // consumer threads
constexpr int kS2Threads = 512;                   // + loader warp (14) + producer warp (15)
constexpr int kS2LoaderWarp = 14, kS2ProdWarp = 15;
constexpr int kS2SlotBytes = kStreamSlotBytes;    // 9216: 32 rows x 288 bytes (Q4_K, two super-blocks)
constexpr int kS2MaxSlots = 24;
constexpr int kS2TileSlots = 4;
constexpr int kS2ZeroBytes = 512;                 // zero page in front of the x region; B operands of idle columns read base + 96
enum : int { PH_EMBED = 2 };

struct Stream2Params {
    MegaParams mp;            // mp.phases: [EMBED, L x (QKV, ATTN, O, GATE/UP, DOWN), HEAD]
    int xr_off;               // dynamic shared memory: zero page at 0, x region at xr_off
    int tpart_off;            // [kS2TileSlots][kS2Cons][2][32] floats
    int desc_off;             // MegaPhase[2]
    int ring_off;
    int n_slots;
    int no_load;
    uint2* ll;                // [grid][2][32] (value, epoch) packets of tile pieces
    unsigned int epoch0;      // packets of this launch carry epoch0 + global phase number + 1
    float* cand_val;          // [grid * kS2Cons] argmax candidates of the vocab head
    int* cand_idx;
};

__device__ __forceinline__ void s2_cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kS2NT) : "memory"); }

// ---------------------------------------------------------------- producer
struct PDesc2 {
    int gemv, n_seg, ept, parts, swiglu, E;
    const void* tm[3];
    int nt[3], cstep[3], bytes[3];
};
__device__ __forceinline__ void pdesc2_load(PDesc2& d, const MegaPhase* P) {
    d.gemv = P->kind == PH_GEMV;
    const MParams* g = &P->gemv;
    d.n_seg = g->n_seg; d.ept = g->s_ept; d.parts = g->s_parts; d.swiglu = g->epi == ME_SWIGLU; d.E = g->s_E;
    const int sC = g->s_C;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const MSeg* sg = &g->seg[s];
        d.tm[s] = sg->tmap;
        d.nt[s] = sg->n_tiles;
        d.cstep[s] = sC * sg->chunk_bytes;
        d.bytes[s] = sg->s_pitch * kMmaRows;
    }
}

__device__ __forceinline__ void s2_producer(const Stream2Params& sp, const SRing& rg, volatile int* s_dead) {
    const MegaParams& mp = sp.mp;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const long long total = (long long)mp.n_tokens * n_run;
    const long long nb = gridDim.x, b = blockIdx.x;
    uint32_t slot = 0, round = 0;
    PDesc2 cur, nxt;
    pdesc2_load(cur, mp.phases);
    int ph_next = 1 % n_run;
    for (long long it = 0; it < total; it++) {
        pdesc2_load(nxt, mp.phases + ph_next);   // in flight while this phase's entries are issued
        if (++ph_next == n_run) ph_next = 0;
        if (cur.gemv) {
            const int e0 = (int)(b * cur.E / nb), e1 = (int)((b + 1) * cur.E / nb);
            if (e1 > e0) {
                if (it < n_run) {
#pragma unroll
                    for (int s = 0; s < 3; s++)
                        if (s < cur.n_seg) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(cur.tm[s]) : "memory");
                }
                const int ept = cur.ept, per_tile = cur.parts * ept;
                int T = e0 / per_tile;
                const int r = e0 - T * per_tile;
                int part = r / ept, ce = r - part * ept;
                for (int e = e0; e < e1; e++) {
                    int s = 0, tile = T;
                    if (cur.swiglu) {
                        s = part;
                    } else {
                        if (cur.n_seg > 1 && tile >= cur.nt[0]) { tile -= cur.nt[0]; s = 1; }
                        if (s == 1 && cur.n_seg > 2 && tile >= cur.nt[1]) { tile -= cur.nt[1]; s = 2; }
                    }
                    const void* tmap = s == 0 ? cur.tm[0] : s == 1 ? cur.tm[1] : cur.tm[2];
                    const int cs = s == 0 ? cur.cstep[0] : s == 1 ? cur.cstep[1] : cur.cstep[2];
                    const int nby = s == 0 ? cur.bytes[0] : s == 1 ? cur.bytes[1] : cur.bytes[2];
                    const int c0 = ((ce * cs) & ~15) >> 2;   // box start, 16-byte aligned, in 4-byte tensor-map elements
                    if (!s_wait(rg.empty + 8u * slot, (round & 1u) ^ 1u, s_dead, mp.err, 1000, round * (uint32_t)rg.n_slots + slot)) return;
                    if (sp.no_load) {
                        mbar_arrive(rg.full + 8u * slot);
                    } else {
                        mbar_arrive_expect_tx(rg.full + 8u * slot, (uint32_t)nby);
                        tma_load_2d(rg.base + slot * (uint32_t)kS2SlotBytes, tmap, c0, tile * kMmaRows, rg.full + 8u * slot);
                    }
                    if (++slot == (uint32_t)rg.n_slots) { slot = 0; round++; }
                    if (++ce == ept) {
                        ce = 0;
                        if (++part == cur.parts) { part = 0; T++; }
                    }
                }
            }
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------- consumer side of one GEMV phase
struct S2Cons {
    uint32_t seq0;     // ring entries this CTA has consumed before this phase
    uint32_t tseq0;    // tiles this CTA has touched before this phase (tile slot = sequence & 3, generation = sequence >> 2)
    float best_v;      // greedy: this lane's best logit so far (vocab head phase)
    int best_i;
};

__device__ __forceinline__ void s2_gemv_cta(const MParams& p, const Stream2Params& sp, uint8_t* smem, const SRing& rg, S2Cons& cs,
                                            uint32_t xfull, uint32_t xpar, int* s_tcnt, volatile unsigned int* s_tdone,
                                            volatile int* s_dead, unsigned int epoch, bool greedy) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int K = p.K;
    const uint32_t sbase = smem_u32(smem);
    const bool swiglu = p.epi == ME_SWIGLU;
    const int ept = p.s_ept, n_parts = p.s_parts, per_tile = n_parts * ept;
    const long long nb = gridDim.x, b = blockIdx.x, E = p.s_E;
    const int e0 = (int)(b * E / nb), e1 = (int)((b + 1) * E / nb), nloc = e1 - e0;
    const int T_first = e0 / per_tile;
    const int n_ltiles = nloc > 0 ? (e1 - 1) / per_tile - T_first + 1 : 0;
    float (*s_tpart)[kS2Cons][2][32] = reinterpret_cast<float (*)[kS2Cons][2][32]>(smem + sp.tpart_off);
    const bool cand = greedy && p.cand;

    const XLayout XL = x_layout(K);
    XSmem sm;
    const uint32_t xb = sbase + (uint32_t)sp.xr_off;
    sm.p0 = xb + XL.p0;
    sm.p1 = xb + XL.p1;
    sm.p2 = xb + XL.p2;
    sm.sx = xb + XL.sx;
    sm.x16 = xb + XL.x16;
    sm.zero = sbase + 96u;

    // logical tile T of the phase -> (segment, tile within the segment)
    auto seg_of = [&](int T, int& s, int& tile) {
        s = 0;
        tile = T;
        if (!swiglu)
            while (s + 1 < p.n_seg && tile >= p.seg[s].n_tiles) { tile -= p.seg[s].n_tiles; s++; }
    };
    // cursor of this warp's pair: local entry index i = pair, pair + 7, ...
    const int pair = warp >> 1, half = warp & 1;
    int i = pair, T = 0, part = 0, ce = 0, s = 0, tile = 0;
    if (i < nloc) {
        const int e = e0 + i;
        T = e / per_tile;
        const int r = e - T * per_tile;
        part = r / ept;
        ce = r - part * ept;
        seg_of(T, s, tile);
    }
    uint32_t q = cs.seq0 + (uint32_t)pair;
    uint32_t slot = q % (uint32_t)rg.n_slots, round = q / (uint32_t)rg.n_slots;
    int type = -1, nb_row = 0, cb = 0;
    uint32_t RS = 0, cbytes = 0;
    LaneB lb{};
    auto load_mat = [&]() {
        const MSeg& wsg = p.seg[swiglu ? part : s];
        RS = (uint32_t)wsg.s_pitch;
        cbytes = (uint32_t)wsg.chunk_bytes;
        nb_row = wsg.nb_row;
        cb = wsg.cb;
        if (wsg.type != type) {
            type = wsg.type;
            lb = (type == T_Q6_K) ? lane_b_q6k(sm, g, t) : (type == T_Q8_0) ? lane_b_q80(sm, g, t) : lane_b_k45(sm, g, t, type == T_Q5_K);
        }
    };
    if (i < nloc) load_mat();
    const int sC = p.s_C, n_chunks = p.chunks;

    // ---- the phase's input: staged by its producer, copied by the loader warp; everything above overlapped the boundary ----
    s_wait(xfull, xpar, s_dead, p.err, 6000, epoch);
    float unscale = 1.0f;
    if (p.norm_w) {   // sum of x^2 from the per-group partial sums, in the same order in every warp and CTA
        const uint32_t ssq = xb + XL.ssq + smem_token();
        float tot = 0.0f;
        for (int k = lane; k < (K >> 5); k += 32) tot += lds_f32(ssq + 4u * (uint32_t)k);
        tot = warp_sum(tot);
        unscale = 1.0f / sqrtf(tot / (float)K + p.eps);
    }
    {   // no shared-memory read of x may be hoisted above the wait (lane tables hold differences: they stay valid)
        const uint32_t tokx = smem_token();
        sm.sx += tokx;
        sm.x16 += tokx;
        sm.zero += tokx;
    }

    // ---- epilogue of a finished tile: lane L owns row tile * 32 + L of segment s (vu: the up row for SwiGLU) ----
    auto epilogue = [&](int es, int etile, float v, float vu) {
        const MSeg& sg = p.seg[es];
        const int j = etile * kMmaRows + lane;
        const bool valid = j < sg.n_rows;
        const float e_bias = (valid && sg.bias) ? sg.bias[j] : 0.0f;
        const float e_res = (valid && p.epi == ME_RESIDUAL) ? __ldcg(p.residual + j) : 0.0f;
        const float e_w = (p.stage_out && p.stage_w && j < p.stage_K) ? p.stage_w[j] : 1.0f;
        v *= unscale;
        float val = v;
        if (swiglu) val = mma_silu(v) * (vu * unscale);
        if (valid) {
            val += e_bias;
            val += e_res;
            if (cand) {   // raw-logit argmax, LAST maximal index wins (src/main.rs:1816-1821)
                if (val > cs.best_v || (val == cs.best_v && j > cs.best_i) || cs.best_i < 0) { cs.best_v = val; cs.best_i = j; }
            } else {
                sg.out[j] = val;
            }
        }
        if (p.stage_out) stage_out32(val, e_w, j, p.stage_K, p.stage_out);
    };

    float ag[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f};
    while (i < nloc) {
        const int c0 = ce * sC, nc = min(sC, n_chunks - c0);
        const uint32_t e00 = (uint32_t)c0 * kMmaChunk;
        const uint32_t doff = ((uint32_t)c0 * cbytes) & 15u;   // the box starts 16-byte aligned (Q6_K: any even residue)
        // parity protocol of stream.cuh: the slot's previous round must have been released before "parity r" of `full`
        // can be trusted (no pair is ever more than n_slots entries ahead of the slowest one: n_slots > kS2Pairs)
        s_wait(rg.empty + 8u * slot, (round & 1u) ^ 1u, s_dead, p.err, 4000 + warp, q);
        s_wait(rg.full + 8u * slot, round & 1u, s_dead, p.err, 2000 + warp, q);
        const uint32_t spb = rg.base + slot * (uint32_t)kS2SlotBytes + doff + smem_token();
        float ua[4] = {0.f, 0.f, 0.f, 0.f};
        switch (type) {
            case T_Q4_K:
                for (int c = 0; c < nc; c++) {
                    if (half == 0) unit2_k45_half<false, 0>(spb + (uint32_t)c * cbytes, RS, e00 + (uint32_t)c * kMmaChunk, sm, lb, g, t, ua);
                    else unit2_k45_half<false, 1>(spb + (uint32_t)c * cbytes, RS, e00 + (uint32_t)c * kMmaChunk, sm, lb, g, t, ua);
                }
                break;
            case T_Q5_K:
                for (int c = 0; c < nc; c++) {
                    if (half == 0) unit2_k45_half<true, 0>(spb + (uint32_t)c * cbytes, RS, e00 + (uint32_t)c * kMmaChunk, sm, lb, g, t, ua);
                    else unit2_k45_half<true, 1>(spb + (uint32_t)c * cbytes, RS, e00 + (uint32_t)c * kMmaChunk, sm, lb, g, t, ua);
                }
                break;
            case T_Q6_K:
                for (int c = 0; c < nc; c++) {
                    const uint32_t a = spb + (uint32_t)c * cbytes, ee = e00 + (uint32_t)c * kMmaChunk;
                    const uint32_t al = (doff + (uint32_t)c * cbytes) & 7u;
                    if (half == 0) {
                        if (al == 0u) unit2_q6k_half<8, 0>(a, RS, ee, sm, lb, g, t, ua);
                        else if (al == 4u) unit2_q6k_half<4, 0>(a, RS, ee, sm, lb, g, t, ua);
                        else unit2_q6k_half<2, 0>(a, RS, ee, sm, lb, g, t, ua);
                    } else {
                        if (al == 0u) unit2_q6k_half<8, 1>(a, RS, ee, sm, lb, g, t, ua);
                        else if (al == 4u) unit2_q6k_half<4, 1>(a, RS, ee, sm, lb, g, t, ua);
                        else unit2_q6k_half<2, 1>(a, RS, ee, sm, lb, g, t, ua);
                    }
                }
                break;
            default:
                for (int c = 0; c < nc; c++) {
                    const uint32_t ee = e00 + (uint32_t)c * kMmaChunk;
                    const int nblk = min(cb, nb_row - (int)(ee >> 5));
                    if (half == 0) unit2_q80_half<0>(spb + (uint32_t)c * cbytes, RS, ee, nblk, sm, lb, g, t, ua);
                    else unit2_q80_half<1>(spb + (uint32_t)c * cbytes, RS, ee, nblk, sm, lb, g, t, ua);
                }
                break;
        }
        pin4(ua);   // every shared-memory read of the entry has completed before the slot is handed back
        __syncwarp();
        if (lane == 0) mbar_arrive(rg.empty + 8u * slot);
        q += (uint32_t)kS2Pairs;
        slot += (uint32_t)kS2Pairs;
        while (slot >= (uint32_t)rg.n_slots) { slot -= (uint32_t)rg.n_slots; round++; }
        if (swiglu && part == 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) au[k] += ua[k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) ag[k] += ua[k];
        }

        // ---- advance the cursor by 7 entries ----
        const int cur_T = T, cur_s = s, cur_tile = tile;
        i += kS2Pairs;
        ce += kS2Pairs;
        bool mat_change = false;
        while (ce >= ept) {
            ce -= ept;
            mat_change = true;
            if (++part == n_parts) { part = 0; T++; }
        }
        if (i < nloc && mat_change) {
            if (T != cur_T) seg_of(T, s, tile);
            load_mat();
        }
        if (i < nloc && T == cur_T) continue;

        // ---- this warp is done with tile cur_T: its 32 row sums go to the tile's slot; the last warp to arrive finishes the tile ----
        float vg = 0.f, vu = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ag[k] += __shfl_xor_sync(0xffffffffu, ag[k], 1);
            ag[k] += __shfl_xor_sync(0xffffffffu, ag[k], 2);
            const float x = __shfl_sync(0xffffffffu, ag[k], 4 * (lane & 7));
            if ((lane >> 3) == k) vg = x;
            ag[k] = 0.f;
        }
        if (swiglu) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                au[k] += __shfl_xor_sync(0xffffffffu, au[k], 1);
                au[k] += __shfl_xor_sync(0xffffffffu, au[k], 2);
                const float x = __shfl_sync(0xffffffffu, au[k], 4 * (lane & 7));
                if ((lane >> 3) == k) vu = x;
                au[k] = 0.f;
            }
        }
        const int t_lo = max(cur_T * per_tile, e0) - e0, t_hi = min((cur_T + 1) * per_tile, e1) - e0;   // local entries of the tile
        const int n_cpairs = min(t_hi - t_lo, kS2Pairs), n_contrib = 2 * n_cpairs;   // both warps of a pair always contribute
        const uint32_t ts = cs.tseq0 + (uint32_t)(cur_T - T_first);
        const int tslot = (int)(ts & (kS2TileSlots - 1));
        const unsigned int gen = ts >> 2;
        {
            {   // the slot's previous tile must have been merged (practically always true: warps are at most a ring apart)
                const long long w0 = clock64();
                while (s_tdone[tslot] != gen) {
                    if (*s_dead) break;
                    if (clock64() - w0 > 2000000000LL) {
                        *s_dead = 1;
                        if (atomicExch(p.err, 5) == 0) { p.err[1] = 7000 + warp; p.err[2] = (int)blockIdx.x; p.err[3] = (int)ts; }
                        break;
                    }
                }
            }
            s_tpart[tslot][warp][0][lane] = vg;
            s_tpart[tslot][warp][1][lane] = vu;
            __syncwarp();
            int old = 0;
            if (lane == 0) {
                __threadfence_block();
                old = atomicAdd(&s_tcnt[tslot], 1);
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old != n_contrib - 1) continue;
            __threadfence_block();
            vg = 0.f;
            vu = 0.f;
            int w = 2 * (t_lo % kS2Pairs);
            for (int k = 0; k < n_contrib; k++) {   // entry order, half 0 before half 1: fixed, whoever arrives last
                vg += s_tpart[tslot][w][0][lane];
                vu += s_tpart[tslot][w][1][lane];
                if (++w == kS2Cons) w = 0;
            }
            __syncwarp();
            if (lane == 0) {
                s_tcnt[tslot] = 0;
                __threadfence_block();
                s_tdone[tslot] = gen + 1u;
            }
        }
        // ---- a tile cut by a CTA boundary: tail pieces are published, the head's CTA collects them ----
        const bool head_local = cur_T * per_tile >= e0, tail_local = (cur_T + 1) * per_tile <= e1;
        if (!head_local) {
            uint2* mine = sp.ll + (size_t)blockIdx.x * 64;
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(mine + lane), "r"(__float_as_uint(vg)), "r"(epoch) : "memory");
            if (swiglu) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(mine + 32 + lane), "r"(__float_as_uint(vu)), "r"(epoch) : "memory");
            continue;
        }
        if (!tail_local) {
            const int h_last = (int)((((long long)(cur_T + 1) * per_tile) * nb - 1) / E);   // CTA of the tile's last entry
            for (int h = (int)blockIdx.x + 1; h <= h_last; h++) {
                if ((long long)h * E / nb == (long long)(h + 1) * E / nb) continue;   // a CTA without entries publishes nothing
                const uint2* theirs = sp.ll + (size_t)h * 64;
                uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
                const long long w0 = clock64();
                for (;;) {
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "l"(theirs + lane) : "memory");
                    if (swiglu) asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "l"(theirs + 32 + lane) : "memory");
                    else b1 = epoch;
                    if (__all_sync(0xffffffffu, a1 == epoch && b1 == epoch)) break;
                    if (*s_dead || clock64() - w0 > 2000000000LL) {
                        *s_dead = 1;
                        if (atomicExch(p.err, 6) == 0) { p.err[1] = 8000 + warp; p.err[2] = (int)blockIdx.x; p.err[3] = h; }
                        break;
                    }
                }
                vg += __uint_as_float(a0);
                vu += __uint_as_float(b0);
            }
        }
        epilogue(cur_s, cur_tile, vg, vu);
    }
    cs.seq0 += (uint32_t)nloc;
    cs.tseq0 += (uint32_t)n_ltiles;
    if (cand) {   // the warp's candidate of this token: larger value, then larger index
        float bv = cs.best_v;
        int bi = cs.best_i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            sp.cand_val[blockIdx.x * kS2Cons + warp] = bv;
            sp.cand_idx[blockIdx.x * kS2Cons + warp] = bi;
        }
        cs.best_v = -INFINITY;
        cs.best_i = -1;
    }
}

// ---------------------------------------------------------------- attention phase (consumer warps)
// Shared memory (floats, from the start of the x region): raw q of the kv head's group [GMAX * HD] | raw k [HD] | raw v [HD]
// (loader's bulk copies) | rotated q [GMAX * HD] | rotated k [HD] | s_m [NW * GMAX] | s_l [NW * GMAX] | s_acc [NW * GMAX * HD];
// the ticket merge of a split context reuses the area from s_m on ([ns][G][HD + 2]).
__host__ __device__ inline size_t attn2_smem_floats(int hd, int gmax, int nw, int n_splits, int G) {
    const size_t head = (size_t)2 * gmax * hd + 3 * (size_t)hd;
    const size_t a = (size_t)2 * nw * gmax + (size_t)nw * gmax * hd, b = (size_t)n_splits * G * (hd + 2);
    return head + (a > b ? a : b);
}

template <int HD, int GMAX, int NW>
__device__ __forceinline__ void attn2_phase(const AttnParams& p, int kv_len, float* xr, uint32_t xfull, uint32_t xpar, volatile int* s_dead,
                                            int* err, unsigned int epoch, unsigned int* s_ticket, const float* s_rope) {
    constexpr int VEC = HD / 32;
    constexpr int NT = NW * 32;
    float* s_qraw = xr;
    float* s_kraw = xr + GMAX * HD;
    float* s_vraw = s_kraw + HD;
    float* s_q = s_vraw + HD;
    float* s_k = s_q + GMAX * HD;
    float* s_m = s_k + HD;                // [warps][GMAX]
    float* s_l = s_m + NW * GMAX;         // [warps][GMAX]
    float* s_acc = s_l + NW * GMAX;       // [warps][GMAX][HD]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.G;
    const int ns = attn_eff_splits(kv_len, p.n_splits, p.min_chunk);
    const int item = blockIdx.x;
    const bool active = item < p.n_kv * ns;
    const int kh = active ? item / ns : 0, split = active ? item - kh * ns : 0;
    int chunk = (kv_len + ns - 1) / ns;
    chunk = (chunk + NW - 1) / NW * NW;
    const int start = split * chunk;
    const int end = min(kv_len, start + chunk);
    const int pos = kv_len - 1;
    const bool own = active && pos >= start && pos < end;   // this split holds the new position (always its last one)
    const int end_g = own ? end - 1 : end;                  // rows [start, end_g) come from the cache in global memory

    const float* kb = p.k_cache + kv_row(p, kh, 0, HD) + lane * VEC;
    const float* vb = p.v_cache + kv_row(p, kh, 0, HD) + lane * VEC;
    const size_t pstride = p.kv_pos_stride ? (size_t)p.kv_pos_stride : (size_t)HD;
    constexpr int UB = GMAX <= 4 ? 2 : 1;   // positions per batch: 128 registers per thread
    constexpr int STEP = NW * UB;
    auto load = [&](int pos0, float (&kr)[UB][VEC], float (&vr)[UB][VEC]) {
#pragma unroll
        for (int u = 0; u < UB; u++) {
            const int pp = pos0 + u * NW;
            const int pc = pp < end_g ? pp : pos0;  // clamp: loads stay in range, result discarded
            if constexpr (VEC == 4) {
                const float4 a = __ldcg(reinterpret_cast<const float4*>(kb + (size_t)pc * pstride));
                const float4 c = __ldcg(reinterpret_cast<const float4*>(vb + (size_t)pc * pstride));
                kr[u][0] = a.x; kr[u][1] = a.y; kr[u][2] = a.z; kr[u][3] = a.w;
                vr[u][0] = c.x; vr[u][1] = c.y; vr[u][2] = c.z; vr[u][3] = c.w;
            } else {
                const float2 a = __ldcg(reinterpret_cast<const float2*>(kb + (size_t)pc * pstride));
                const float2 c = __ldcg(reinterpret_cast<const float2*>(vb + (size_t)pc * pstride));
                kr[u][0] = a.x; kr[u][1] = a.y;
                vr[u][0] = c.x; vr[u][1] = c.y;
            }
        }
    };
    // rows of earlier positions do not depend on this token: request the first batch before waiting for the boundary
    float kA[UB][VEC], vA[UB][VEC];
    int pos0 = start + warp;
    if (active && pos0 < end_g) load(pos0, kA, vA);

    s_wait(xfull, xpar, s_dead, err, 6100, epoch);
    if (!active) return;   // the whole CTA (all consumer threads) leaves together

    {   // Backend::rope (cpu/ops.rs:1216-1337) on the raw projections, cache write (layers.rs:580-600)
        const int half = HD / 2;
        for (int idx = tid; idx < G * half; idx += NT) {
            const int gq = idx / half, pi = idx - gq * half;
            const int i0 = p.neox ? pi : 2 * pi, i1 = p.neox ? pi + half : 2 * pi + 1;
            const float c = s_rope[pi], sn = s_rope[half + pi];
            const float x0 = s_qraw[gq * HD + i0], x1 = s_qraw[gq * HD + i1];
            s_q[gq * HD + i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, sn));
            s_q[gq * HD + i1] = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
        }
        if (own) {
            float* ko = const_cast<float*>(p.k_cache) + kv_row(p, kh, pos, HD);
            float* vo = const_cast<float*>(p.v_cache) + kv_row(p, kh, pos, HD);
            for (int pi = tid; pi < half; pi += NT) {
                const int i0 = p.neox ? pi : 2 * pi, i1 = p.neox ? pi + half : 2 * pi + 1;
                const float c = s_rope[pi], sn = s_rope[half + pi];
                const float k0 = s_kraw[i0], k1 = s_kraw[i1];
                const float r0 = __fsub_rn(__fmul_rn(k0, c), __fmul_rn(k1, sn)), r1 = __fadd_rn(__fmul_rn(k0, sn), __fmul_rn(k1, c));
                s_k[i0] = r0;
                s_k[i1] = r1;
                ko[i0] = r0;
                ko[i1] = r1;
            }
            for (int d = tid; d < HD; d += NT) vo[d] = s_vraw[d];
        }
    }
    s2_cons_sync();

    float q[GMAX][VEC], acc[GMAX][VEC], m[GMAX], l[GMAX];
#pragma unroll
    for (int gq = 0; gq < GMAX; gq++) {
        m[gq] = -INFINITY;
        l[gq] = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            acc[gq][v] = 0.0f;
            q[gq][v] = (gq < G) ? s_q[gq * HD + lane * VEC + v] : 0.0f;
        }
    }
    auto compute = [&](int p0, int lim, const float (&kr)[UB][VEC], const float (&vr)[UB][VEC]) {
        float s[UB][GMAX];
#pragma unroll
        for (int u = 0; u < UB; u++)
#pragma unroll
            for (int gq = 0; gq < GMAX; gq++) {
                float d = 0.0f;
#pragma unroll
                for (int v = 0; v < VEC; v++) d = fmaf(q[gq][v], kr[u][v], d);
                s[u][gq] = d;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < UB; u++)
#pragma unroll
                for (int gq = 0; gq < GMAX; gq++) s[u][gq] += __shfl_xor_sync(0xffffffffu, s[u][gq], o);
#pragma unroll
        for (int gq = 0; gq < GMAX; gq++) {
            if (gq < G) {
                float mb = -INFINITY;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    s[u][gq] = (p0 + u * NW < lim) ? s[u][gq] * p.scale : -INFINITY;   // warp-uniform
                    mb = fmaxf(mb, s[u][gq]);
                }
                const float mn = fmaxf(m[gq], mb);   // finite: the first position of a batch is always valid
                const float corr = (m[gq] == -INFINITY) ? 0.0f : expf(m[gq] - mn);
                float w[UB], ws = 0.0f;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    w[u] = (s[u][gq] == -INFINITY) ? 0.0f : expf(s[u][gq] - mn);
                    ws += w[u];
                }
                l[gq] = l[gq] * corr + ws;
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    float a = acc[gq][v] * corr;
#pragma unroll
                    for (int u = 0; u < UB; u++) a = fmaf(w[u], vr[u][v], a);
                    acc[gq][v] = a;
                }
                m[gq] = mn;
            }
        }
    };
    {
        float kB[UB][VEC], vB[UB][VEC];
        while (pos0 < end_g) {
            if (pos0 + STEP < end_g) load(pos0 + STEP, kB, vB);
            compute(pos0, end_g, kA, vA);
            pos0 += STEP;
            if (pos0 >= end_g) break;
            if (pos0 + STEP < end_g) load(pos0 + STEP, kA, vA);
            compute(pos0, end_g, kB, vB);
            pos0 += STEP;
        }
    }
    if (own && warp == (pos - start) % NW) {   // the new position, from shared memory (one valid row in a batch of UB)
        float kr[UB][VEC], vr[UB][VEC];
#pragma unroll
        for (int u = 0; u < UB; u++)
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                kr[u][v] = s_k[lane * VEC + v];
                vr[u][v] = s_vraw[lane * VEC + v];
            }
        compute(0, 1, kr, vr);
    }

    // ---- combine the warps of this CTA ----
#pragma unroll
    for (int gq = 0; gq < GMAX; gq++) {
        if (gq < G) {
            if (lane == 0) {
                s_m[warp * GMAX + gq] = m[gq];
                s_l[warp * GMAX + gq] = l[gq];
            }
#pragma unroll
            for (int v = 0; v < VEC; v++) s_acc[(warp * GMAX + gq) * HD + lane * VEC + v] = acc[gq][v];
        }
    }
    s2_cons_sync();
    const int part_stride = HD + 2;
    float* my_part = p.part + ((size_t)(kh * p.n_splits + split) * G) * part_stride;
    if (tid < G) {
        const int gq = tid;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; w++) M = fmaxf(M, s_m[w * GMAX + gq]);
        float L = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const float mw = s_m[w * GMAX + gq];
            const float c = (mw == -INFINITY) ? 0.0f : expf(mw - M);
            L += s_l[w * GMAX + gq] * c;
            s_m[w * GMAX + gq] = c;
        }
        s_l[0 * GMAX + gq] = M;
        s_l[1 * GMAX + gq] = L;
    }
    s2_cons_sync();
    // a warp finishes 32 consecutive elements of the output vector at a time (the staged form needs the whole group)
    for (int grp = warp; grp < G * (HD / 32); grp += NW) {
        const int gq = grp / (HD / 32), d = (grp - gq * (HD / 32)) * 32 + lane;
        const float M = s_l[0 * GMAX + gq], L = s_l[1 * GMAX + gq];
        float A = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; w++) A += s_acc[(w * GMAX + gq) * HD + d] * s_m[w * GMAX + gq];
        if (ns == 1) {
            const float o = A / L;
            p.out[(kh * G + gq) * HD + d] = o;
            if (p.stage_out) attn_stage_out(o, (kh * G + gq) * HD + d, p.stage_K, p.stage_out);
        } else {
            my_part[gq * part_stride + d] = A;
            if (d == 0) {
                my_part[gq * part_stride + HD] = M;
                my_part[gq * part_stride + HD + 1] = L;
            }
        }
    }
    if (ns == 1) return;

    // ---- last CTA of this kv head merges the splits (fixed order) ----
    s2_cons_sync();
    if (tid == 0) {
        unsigned int tk;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(tk) : "l"(p.tickets + kh) : "memory");
        *s_ticket = tk;
    }
    s2_cons_sync();
    if (*s_ticket != (unsigned)(ns - 1)) return;
    const float* parts = p.part + (size_t)kh * p.n_splits * G * part_stride;
    const int n_part = ns * G * part_stride;
    float* s_p = s_m;                            // [ns][G][HD + 2]
    for (int i0 = tid; i0 < n_part; i0 += NT * 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (i0 + k * NT < n_part) ? __ldcg(parts + i0 + k * NT) : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (i0 + k * NT < n_part) s_p[i0 + k * NT] = v[k];
    }
    s2_cons_sync();
    if (tid < G) {
        const int gq = tid;
        float M = -INFINITY;
        for (int sI = 0; sI < ns; sI++) M = fmaxf(M, s_p[(sI * G + gq) * part_stride + HD]);
        float L = 0.0f;
        for (int sI = 0; sI < ns; sI++) {
            float* ps = s_p + (sI * G + gq) * part_stride;
            const float ms = ps[HD];
            const float c = (ms == -INFINITY) ? 0.0f : expf(ms - M);
            L += ps[HD + 1] * c;
            ps[HD] = c;
        }
        s_p[(0 * G + gq) * part_stride + HD + 1] = L;
    }
    s2_cons_sync();
    for (int grp = warp; grp < G * (HD / 32); grp += NW) {
        const int gq = grp / (HD / 32), d = (grp - gq * (HD / 32)) * 32 + lane;
        const float L = s_p[(0 * G + gq) * part_stride + HD + 1];
        float A = 0.0f;
        for (int sI = 0; sI < ns; sI++) {
            const float* ps = s_p + (sI * G + gq) * part_stride;
            A += ps[d] * ps[HD];
        }
        const float o = A / L;
        p.out[(kh * G + gq) * HD + d] = o;
        if (p.stage_out) attn_stage_out(o, (kh * G + gq) * HD + d, p.stage_K, p.stage_out);
    }
    if (tid == 0) p.tickets[kh] = 0;  // ready for the next layer / launch
}

// ---------------------------------------------------------------- CTA 0: pick (greedy) + embedding row + its staged form
__device__ __forceinline__ void s2_pick(const Stream2Params& sp, float* s_av, int* s_ai, int* s_tok) {
    const MegaParams& mp = sp.mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)gridDim.x * kS2Cons;
    float best = -INFINITY;
    int bi = -1;
    for (int c = tid; c < n; c += kS2NT) {
        const float v = __ldcg(sp.cand_val + c);
        const int i = __ldcg(sp.cand_idx + c);
        if (i >= 0 && (bi < 0 || v > best || (v == best && i > bi))) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_av[warp] = best; s_ai[warp] = bi; }
    s2_cons_sync();
    if (tid == 0) {
        for (int w = 1; w < kS2Cons; w++)
            if (s_ai[w] >= 0 && (bi < 0 || s_av[w] > best || (s_av[w] == best && s_ai[w] > bi))) { best = s_av[w]; bi = s_ai[w]; }
        mp.st->token = bi;
        const int gcount = mp.st->n_generated;
        if (gcount < mp.max_generated) mp.generated[gcount] = bi;
        mp.st->n_generated = gcount + 1;
        *s_tok = bi;
    }
    s2_cons_sync();
}

__device__ __forceinline__ void s2_embed(const MParams& p, const Stream2Params& sp, bool pick, float* s_av, int* s_ai, int* s_tok) {
    const MegaParams& mp = sp.mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (pick) {
        s2_pick(sp, s_av, s_ai, s_tok);
    } else {
        if (tid == 0) *s_tok = __ldcg(&mp.st->token);
        s2_cons_sync();
    }
    int token = *s_tok;
    token = min(max(token, 0), mp.vocab - 1);
    if (tid == 0) {
        const int pn = mp.st->pos_next;
        mp.st->pos_cur = pn;
        mp.st->pos_next = pn + 1;
    }
    // LlamaModel::forward (model/llama.rs:293-306): row `token` of token_embd, dequantised bit-exactly
    const int be = type_block_elems(mp.embd_type), bb = type_block_bytes(mp.embd_type);
    const uint8_t* row = mp.embd + (long long)token * mp.embd_row_bytes;
    for (int grp = warp; grp < (mp.hidden >> 5); grp += kS2Cons) {
        const int i = grp * 32 + lane;
        const int blk = i / be;
        const float v = dequant_elem(mp.embd_type, row + (long long)blk * bb, i - blk * be);
        mp.h[i] = v;
        stage_out32(v, p.stage_w ? p.stage_w[i] : 1.0f, i, p.stage_K, p.stage_out);
    }
}

// ---------------------------------------------------------------- loader warp: the phase boundary
__device__ __forceinline__ void s2_loader(const Stream2Params& sp, uint8_t* smem, uint32_t xfull, uint32_t done, int* s_pos, float* s_rope,
                                          int hd, volatile int* s_dead) {
    const MegaParams& mp = sp.mp;
    const int lane = threadIdx.x & 31;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const long long total = (long long)mp.n_tokens * n_run;
    const bool final_pick = mp.mode == MEGA_GREEDY;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t xr = sbase + (uint32_t)sp.xr_off;
    const MegaPhase* s_desc = reinterpret_cast<const MegaPhase*>(smem + sp.desc_off);
    unsigned int target = 0;
    int kv_len = 1;
    int ph = 0;
    for (long long gb = 0; gb < total + (final_pick ? 1 : 0); gb++) {
        const bool last = gb == total;     // the boundary after the last vocab head (greedy: CTA 0 picks the last token)
        bool ok = true;
        uint32_t tx = 0;
        if (lane == 0) {
            if (gb > 0) {
                ok = s_wait(done, (uint32_t)((gb - 1) & 1), s_dead, mp.err, 9000, (uint32_t)gb);
                s_wait(xfull, (uint32_t)((gb - 1) & 1), s_dead, mp.err, 9100, (uint32_t)gb);   // s_desc[gb & 1] has landed
            }
            if (!last && gb + 1 < total) {   // descriptor of the NEXT phase: constant data, issued before the barrier is polled
                int nph = ph + 1;
                if (nph == n_run) nph = 0;
                tx += (uint32_t)sizeof(MegaPhase);
                bulk_g2s(sbase + (uint32_t)sp.desc_off + (uint32_t)(((gb + 1) & 1) * sizeof(MegaPhase)), mp.phases + nph,
                         (uint32_t)sizeof(MegaPhase), xfull);
            }
            if (gb > 0) {   // grid barrier: everything every CTA wrote in phase gb - 1 is visible after this
                target += gridDim.x;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(mp.bar) : "memory");
                unsigned int v = 0;
                const long long t0 = clock64();
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(mp.bar) : "memory");
                    if (v >= target) break;
                    if (*s_dead || clock64() - t0 > 3000000000LL) {
                        *s_dead = 1;
                        if (atomicExch(mp.err, 2) == 0) { mp.err[1] = 3000; mp.err[2] = (int)blockIdx.x; mp.err[3] = (int)target; }
                        break;
                    }
                }
                if (mp.dbg && blockIdx.x == 0) mp.dbg[last ? n_run : ph] = gtimer();   // boundary BEFORE phase ph of the current token
                asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes of other CTAs -> this thread's bulk copies
            }
        }
        if (!last) {
            const MegaPhase& cur = s_desc[gb & 1];
            if (ph == 1) {   // first phase after the embedding: the token's position (written by CTA 0 in the EMBED phase)
                int pc = 0;
                if (lane == 0) pc = __ldcg(&mp.st->pos_cur);
                pc = __shfl_sync(0xffffffffu, pc, 0);
                kv_len = pc + 1;
                if (lane == 0) *s_pos = pc;
            }
            if (lane == 0) {
                if (cur.kind == PH_GEMV) {
                    const uint32_t nbx = (uint32_t)x_staged_bytes(cur.gemv.K);
                    tx += nbx;
                    bulk_g2s(xr, cur.gemv.x_staged, nbx, xfull);
                } else if (cur.kind == PH_ATTN) {
                    const AttnParams& ap = cur.attn;
                    const int ns = attn_eff_splits(kv_len, ap.n_splits, ap.min_chunk);
                    const int item = blockIdx.x;
                    if (item < ap.n_kv * ns) {
                        const int kh = item / ns;
                        const int gmax = ap.G <= 4 ? 4 : 8;
                        const uint32_t qb = (uint32_t)(ap.G * hd * 4), rb = (uint32_t)(hd * 4);
                        tx += qb + 2u * rb;
                        bulk_g2s(xr, ap.qkv_raw + (size_t)kh * ap.G * hd, qb, xfull);
                        bulk_g2s(xr + (uint32_t)(gmax * hd * 4), ap.qkv_raw + (size_t)ap.n_heads * hd + (size_t)kh * hd, rb, xfull);
                        bulk_g2s(xr + (uint32_t)((gmax + 1) * hd * 4), ap.qkv_raw + (size_t)(ap.n_heads + ap.n_kv) * hd + (size_t)kh * hd, rb, xfull);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(xfull, tx);
        if (!last && ph == 1) {   // the rotation angles of this token (Backend::rope, cpu/ops.rs:1216-1337): needed from phase 2 on
            const MegaPhase& nxt = s_desc[gb & 1];   // (any descriptor would do: freq / rope_scale live in the ATTN phases)
            (void)nxt;
            const AttnParams& ap = reinterpret_cast<const MegaPhase*>(mp.phases + 2)->attn;
            const float position = (float)(kv_len - 1) / ap.rope_scale;
            for (int pi = lane; pi < hd / 2; pi += 32) {
                const float theta = position * ap.freq[pi];
                s_rope[pi] = cosf(theta);
                s_rope[hd / 2 + pi] = sinf(theta);
            }
            __syncwarp();
        }
        (void)ok;
        if (++ph == n_run) ph = 0;
    }
}

// ---------------------------------------------------------------- the kernel
template <int HD, int GMAX>
__global__ void __launch_bounds__(kS2Threads, 1) stream2_decode_kernel(const __grid_constant__ Stream2Params sp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bars[2 * kS2MaxSlots + 2];
    __shared__ int s_tcnt[kS2TileSlots];
    __shared__ unsigned int s_tdone[kS2TileSlots];
    __shared__ unsigned int s_ticket;
    __shared__ int s_dead;
    __shared__ int s_pos;
    __shared__ int s_tok;
    __shared__ float s_rope[HD];
    __shared__ float s_av[kS2Cons];
    __shared__ int s_ai[kS2Cons];

    const MegaParams& mp = sp.mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SRing rg;
    rg.base = smem_u32(smem) + (uint32_t)sp.ring_off;
    rg.full = smem_u32(s_bars);
    rg.empty = rg.full + 8u * (uint32_t)sp.n_slots;
    rg.n_slots = sp.n_slots;
    const uint32_t xfull = rg.full + 16u * (uint32_t)kS2MaxSlots, done = xfull + 8u;
    if (tid == 0) {
        for (int i = 0; i < sp.n_slots; i++) {
            mbar_init(rg.full + 8u * i, 1);
            mbar_init(rg.empty + 8u * i, 2);   // both warps of the consuming pair hand the slot back
        }
        mbar_init(xfull, 1);
        mbar_init(done, kS2Cons);
        s_dead = 0;
        s_pos = 0;
        for (int i = 0; i < kS2TileSlots; i++) { s_tcnt[i] = 0; s_tdone[i] = 0u; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < kS2ZeroBytes / 4) reinterpret_cast<uint32_t*>(smem)[tid] = 0u;
    {   // descriptor of phase 0
        const uint32_t* src = reinterpret_cast<const uint32_t*>(mp.phases);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem + sp.desc_off);
        for (int i = tid; i < (int)(sizeof(MegaPhase) / 4); i += kS2Threads) dst[i] = src[i];
    }
    __syncthreads();   // the only CTA-wide barrier

    if (warp == kS2ProdWarp) {
        if (lane == 0) s2_producer(sp, rg, &s_dead);
        s_drain(&s_dead);
        return;
    }
    if (warp == kS2LoaderWarp) {
        s2_loader(sp, smem, xfull, done, &s_pos, s_rope, HD, &s_dead);
        s_drain(&s_dead);
        return;
    }

    // ---------------- consumer warps ----------------
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const bool greedy = mp.mode == MEGA_GREEDY;
    const MegaPhase* s_desc = reinterpret_cast<const MegaPhase*>(smem + sp.desc_off);
    S2Cons cs{0u, 0u, -INFINITY, -1};
    long long gph = 0;
    for (int tok = 0; tok < mp.n_tokens; tok++) {
        for (int ph = 0; ph < n_run; ph++, gph++) {
            const MegaPhase& cur = s_desc[gph & 1];
            const uint32_t xpar = (uint32_t)(gph & 1);
            const unsigned int epoch = sp.epoch0 + (unsigned int)gph + 1u;
            if (cur.kind == PH_GEMV) {
                s2_gemv_cta(cur.gemv, sp, smem, rg, cs, xfull, xpar, s_tcnt, s_tdone, &s_dead, epoch, greedy);
            } else if (cur.kind == PH_ATTN) {
                attn2_phase<HD, GMAX, kS2Cons>(cur.attn, s_pos + 1, reinterpret_cast<float*>(smem + sp.xr_off), xfull, xpar, &s_dead, mp.err,
                                               epoch, &s_ticket, s_rope);
            } else {
                s_wait(xfull, xpar, &s_dead, mp.err, 6200, epoch);
                if (blockIdx.x == 0) s2_embed(cur.gemv, sp, greedy && tok > 0, s_av, s_ai, &s_tok);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(done);
        }
    }
    if (greedy) {   // the last token's pick, after the boundary that follows its vocab head
        s_wait(xfull, (uint32_t)(gph & 1), &s_dead, mp.err, 6300, (uint32_t)gph);
        if (blockIdx.x == 0) s2_pick(sp, s_av, s_ai, &s_tok);
    }
    s_drain(&s_dead);
}

}  // namespace b200
