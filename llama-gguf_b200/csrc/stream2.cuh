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

#ifndef B200_S2_CONS
#define B200_S2_CONS 8
#endif
constexpr int kS2Cons = B200_S2_CONS;             // consumer warps (14 at 128 registers; 12: warpgroups 0-2 take kS2ConsRegs registers with setmaxnreg)
constexpr int kS2NT = kS2Cons * 32;               // consumer threads
// + the service warpgroup: loader warp, producer warp, two idle warps.  8 consumers: 256 x 200 + 128 x 96 = 63488 registers
// (round 1's fully unrolled unit kernels: few fat warps); 12: 384 x 144 + 128 x 72 = 64512; 14: no setmaxnreg, 128 each
constexpr int kS2Threads = kS2Cons == 8 ? 384 : 512;
// producer warps (the rest of the service warpgroup).  n_slots must be a multiple of it as well: a slot is then always filled by the
// same producer, in order, so its wait on `empty` by phase parity can never see a stale phase
constexpr int kS2Prods = kS2Cons == 8 ? 2 : kS2Cons == 12 ? 3 : 1;
constexpr bool kS2Regs = kS2Cons != 14;
constexpr int kS2ConsRegs = kS2Cons == 8 ? 200 : 144, kS2ServRegs = kS2Cons == 8 ? 96 : 72;   // (inc must fit in what dec released: 256 x 32 <= 128 x 72)
constexpr int kS2LoaderWarp = kS2Cons, kS2ProdWarp = kS2Cons + 1;
constexpr int kS2SlotBytes = kStreamSlotBytes;    // 9216: 32 rows x 288 bytes (Q4_K, two super-blocks)
constexpr int kS2MaxSlots = 24;
constexpr int kS2TileSlots = 4;
constexpr int kS2ZeroBytes = 512;                 // zero page in front of the x region; B operands of idle columns read base + 96
enum : int { PH_EMBED = 2, PH_REDUCE = 3 };   // REDUCE (tensor parallel): sum of the ranks' partial vectors + residual -> f32 vector + staged form

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
    int tp_per_token;         // tensor parallel: cross-GPU exchanges per token E (2 per layer, + 1 for the greedy pick); exchange ids of a launch:
                              // token t's k-th row-parallel exchange = t E + k, the pick that opens token t >= 1 = t E, the final pick = n_tokens E
    float* cand_val;          // [grid] argmax candidates of the vocab head (one per CTA)
    int* cand_idx;
};

__device__ __forceinline__ void s2_cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kS2NT) : "memory"); }

// ---------------------------------------------------------------- producer
struct PDesc2 {
    int gemv, n_seg, ept, parts, swiglu, E;
    unsigned long long* dbg;
    const void* tm[3];
    int nt[3], cstep[3], bytes[3];
};
__device__ __forceinline__ void pdesc2_load(PDesc2& d, const MegaPhase* P) {
    d.gemv = P->kind == PH_GEMV;
    const MParams* g = &P->gemv;
    d.dbg = g->dbg;
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

// One issuing thread is latency-bound: a try_wait that succeeds at once, expect_tx, the operand moves to uniform registers and the
// UTMALDG issue take ~400 clocks per entry (measured: producer stamps of b200_debug_mega_phase), i.e. 23 B/clk for a 9 KB entry --
// the HBM rate per SM with no slack, and the reason round 1's gate/up phase never went below 19 us.  So the ring is fed by
// kS2Prods producer warps; producer `pidx` issues the entries whose global sequence number is pidx modulo kS2Prods (the slot of
// an entry is its sequence number modulo n_slots, whoever issues it).
__device__ __forceinline__ void s2_producer(const Stream2Params& sp, const SRing& rg, volatile int* s_dead, int pidx) {
    const MegaParams& mp = sp.mp;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const long long total = (long long)mp.n_tokens * n_run;
    const long long nb = gridDim.x, b = blockIdx.x;
    uint32_t Q0 = 0;                       // entries of this CTA before the current phase
    PDesc2 cur, nxt;
    pdesc2_load(cur, mp.phases);
    int ph_next = 1 % n_run;
    for (long long it = 0; it < total; it++) {
        pdesc2_load(nxt, mp.phases + ph_next);   // in flight while this phase's entries are issued
        if (++ph_next == n_run) ph_next = 0;
        if (cur.gemv) {
            const int e0 = (int)(b * cur.E / nb), e1 = (int)((b + 1) * cur.E / nb), nloc = e1 - e0;
            int k = (int)(((uint32_t)pidx + (uint32_t)kS2Prods * 1024u - Q0 % (uint32_t)kS2Prods) % (uint32_t)kS2Prods);   // first local entry of this producer
            if (k < nloc) {
                if (it < n_run) {
#pragma unroll
                    for (int s = 0; s < 3; s++)
                        if (s < cur.n_seg) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(cur.tm[s]) : "memory");
                }
                const int ept = cur.ept, per_tile = cur.parts * ept;
                int T = (e0 + k) / per_tile;
                const int r = (e0 + k) - T * per_tile;
                int part = r / ept, ce = r - part * ept;
                uint32_t slot = (Q0 + (uint32_t)k) % (uint32_t)rg.n_slots, round = (Q0 + (uint32_t)k) / (uint32_t)rg.n_slots;
                long long pw = 0;
                const long long pt0 = clock64();
                int n_mine = 0;
                // what one entry needs: tensor map, box coordinates, bytes, slot; then the cursor moves on by kS2Prods entries
                auto next_entry = [&](const void*& tmap, int& c0, int& c1, int& nby, uint32_t& sl, uint32_t& par) {
                    int s = 0, tile = T;
                    if (cur.swiglu) {
                        s = part;
                    } else {
                        if (cur.n_seg > 1 && tile >= cur.nt[0]) { tile -= cur.nt[0]; s = 1; }
                        if (s == 1 && cur.n_seg > 2 && tile >= cur.nt[1]) { tile -= cur.nt[1]; s = 2; }
                    }
                    tmap = s == 0 ? cur.tm[0] : s == 1 ? cur.tm[1] : cur.tm[2];
                    const int cs = s == 0 ? cur.cstep[0] : s == 1 ? cur.cstep[1] : cur.cstep[2];
                    nby = s == 0 ? cur.bytes[0] : s == 1 ? cur.bytes[1] : cur.bytes[2];
                    c0 = ((ce * cs) & ~15) >> 2;   // box start, 16-byte aligned, in 4-byte tensor-map elements
                    c1 = tile * kMmaRows;
                    sl = slot;
                    par = (round & 1u) ^ 1u;
                    k += kS2Prods;
                    ce += kS2Prods;
                    while (ce >= ept) {
                        ce -= ept;
                        if (++part == cur.parts) { part = 0; T++; }
                    }
                    slot += (uint32_t)kS2Prods;
                    if (slot >= (uint32_t)rg.n_slots) { slot -= (uint32_t)rg.n_slots; round++; }
                    n_mine++;
                };
                // two entries per iteration: their try_waits are issued back to back and overlap
                while (k < nloc) {
                    const void *tmA, *tmB = nullptr;
                    int a0, a1, an, b0 = 0, b1 = 0, bn = 0;
                    uint32_t slotA, parA, slotB = 0, parB = 0;
                    next_entry(tmA, a0, a1, an, slotA, parA);
                    const bool two = k < nloc;
                    if (two) next_entry(tmB, b0, b1, bn, slotB, parB);
                    if (cur.dbg) pw -= clock64();
                    const bool okA = mbar_try_wait(rg.empty + 8u * slotA, parA);
                    const bool okB = two ? mbar_try_wait(rg.empty + 8u * slotB, parB) : true;
                    if (!okA && !s_wait(rg.empty + 8u * slotA, parA, s_dead, mp.err, 1000, slotA)) return;
                    if (cur.dbg) pw += clock64();
                    if (sp.no_load) {
                        mbar_arrive(rg.full + 8u * slotA);
                    } else {
                        mbar_arrive_expect_tx(rg.full + 8u * slotA, (uint32_t)an);
                        tma_load_2d(rg.base + slotA * (uint32_t)kS2SlotBytes, tmA, a0, a1, rg.full + 8u * slotA);
                    }
                    if (two) {
                        if (cur.dbg) pw -= clock64();
                        if (!okB && !s_wait(rg.empty + 8u * slotB, parB, s_dead, mp.err, 1001, slotB)) return;
                        if (cur.dbg) pw += clock64();
                        if (sp.no_load) {
                            mbar_arrive(rg.full + 8u * slotB);
                        } else {
                            mbar_arrive_expect_tx(rg.full + 8u * slotB, (uint32_t)bn);
                            tma_load_2d(rg.base + slotB * (uint32_t)kS2SlotBytes, tmB, b0, b1, rg.full + 8u * slotB);
                        }
                    }
                }
                if (cur.dbg && pidx == 0) {   // (debug) producer 0's row of the phase's stamp buffer: first issue, last issue, clocks waiting for slots, entries
                    unsigned long long* d = cur.dbg + ((size_t)blockIdx.x * 16 + kS2ProdWarp) * 8;
                    d[0] = (unsigned long long)pt0;
                    d[1] = (unsigned long long)clock64();
                    d[2] = (unsigned long long)pw;
                    d[3] = (unsigned long long)n_mine;
                }
            }
            Q0 += (uint32_t)nloc;
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------- consumer side of one GEMV phase
// Register discipline (128 per thread, no L1 to spill into: the whole carve-out is shared memory): the only state that
// lives across a unit kernel is the cursor (i, T, part, ce), the ring position (slot, round), one accumulator pair, three
// lane-table words and a few phase constants; everything else is re-read from the phase descriptor in shared memory.
// A consumer warp never touches global memory: merge, cross-CTA packets and the epilogue belong to the loader warp.
// which form of the unit kernels the consumers run (units2.cuh): 5 = K-half outer (B operands of one half live), 4 = row-block outer
#ifndef B200_S2_UNITS
#define B200_S2_UNITS 4
#endif
#define S2_CAT2(a, b) a##b
#define S2_CAT(a, b) S2_CAT2(a, b)
#define S2_UNIT(name) S2_CAT(S2_CAT(unit, B200_S2_UNITS), _##name)
struct S2Cons {
    uint32_t seq0;     // ring entries this CTA has consumed before this phase
    uint32_t tseq0;    // tiles this CTA has touched before this phase (tile slot = sequence & 3, generation = sequence >> 2)
};
struct S2Deal {
    int e0, nloc, T_first, n_ltiles;
};
__device__ __forceinline__ S2Deal s2_deal(const MParams& p) {
    const long long nb = gridDim.x, b = blockIdx.x, E = p.s_E;
    const int per_tile = p.s_parts * p.s_ept;
    S2Deal d;
    d.e0 = (int)(b * E / nb);
    d.nloc = (int)((b + 1) * E / nb) - d.e0;
    d.T_first = d.e0 / per_tile;
    d.n_ltiles = d.nloc > 0 ? (d.e0 + d.nloc - 1) / per_tile - d.T_first + 1 : 0;
    return d;
}
__device__ __forceinline__ void pin2(float (&a)[2]) { asm volatile("" : "+f"(a[0]), "+f"(a[1])::"memory"); }
// debug timeline (b200_debug_mega_phase): SM-clock stamps of lane 0 of every warp.  Consumers: 0 phase entered, 1 x landed, 2 first
// entry landed, 3 last entry computed, 4 phase left.  Loader: 0 consumers done, 1 arrived at the grid barrier, 2 barrier open, 3 x
// issued, 4 first tile merged, 5 last tile's epilogue done
#define S2_STAMP(DBG, k) do { if ((DBG) && (threadIdx.x & 31) == 0) (DBG)[((size_t)blockIdx.x * 16 + (threadIdx.x >> 5)) * 8 + (k)] = (unsigned long long)clock64(); } while (0)

// Tensor parallel, input of a GEMV that follows a row-parallel one: the all-reduce is finished by the consumer warps of EVERY CTA of every
// rank, identically -- x[j] = sum over ranks (rank order) of the partial vectors the ranks left in this GPU's buffer + the residual, staged
// straight into shared memory (the bytes a REDUCE phase would have written to global memory and the loader copied back: one grid boundary
// less per exchange).  CTA b also stores the groups b, b + grid, ... of the f32 vector (a later residual).  Not inlined: the single-GPU
// register allocation of s2_gemv_cta must not see this code (measured: 1 % slower inlined).
__device__ __noinline__ void s2_fold_stage(const MParams& p, uint8_t* xs, float* s_red, uint32_t xstg) {
    const int tid = threadIdx.x;   // consumer threads only: 0 .. kS2NT - 1
    const XLayout L = x_layout(p.K);
    float ss = 0.0f;
    // four float4 per thread and batch, every load of a batch in flight at once (K = 4096: one batch); the 8 lanes that hold a group of 32
    // agree on its scale with three shuffles (split_store4, the staging code of the first megakernel)
    for (int e0 = tid * 4; e0 < p.K; e0 += kS2NT * 16) {
        float4 v[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e0 + i * kS2NT * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            w[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (e < p.K) {
                v[i] = __ldcg(reinterpret_cast<const float4*>(p.xsum + e));
                for (int r = 1; r < p.n_sum; r++) {
                    const float4 b = __ldcg(reinterpret_cast<const float4*>(p.xsum + (size_t)r * p.sum_stride + e));
                    v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
                }
                if (p.x_res) {
                    const float4 b = __ldcg(reinterpret_cast<const float4*>(p.x_res + e));
                    v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
                }
                if (p.norm_w) w[i] = *reinterpret_cast<const float4*>(p.norm_w + e);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e0 + i * kS2NT * 4;
            if (e < p.K) {
                if (p.x_full_out && (e >> 5) % (int)gridDim.x == (int)blockIdx.x) *reinterpret_cast<float4*>(p.x_full_out + e) = v[i];
                ss += split_store4(v[i], w[i], e, xs, L, __activemask());
            }
        }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_red[tid >> 5] = ss;   // sum of x^2 (before the norm weight), one partial per warp: the loader adds them (1 / rms)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (the region is the target of later bulk copies)
    s2_cons_sync();
    if (tid == 0) mbar_arrive(xstg);
}

__device__ __forceinline__ void s2_gemv_cta(const MParams& p, const Stream2Params& sp, uint8_t* smem, const SRing& rg, S2Cons& cs,
                                            uint32_t xfull, uint32_t xpar, uint32_t xstg, float* s_red, int* s_tcnt, volatile unsigned int* s_tdone,
                                            volatile int* s_dead, unsigned int epoch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t sbase = smem_u32(smem);
    const bool swiglu = p.epi == ME_SWIGLU;
    const int ept = p.s_ept, per_tile = p.s_parts * ept;
    const S2Deal dl = s2_deal(p);
    const int e0 = dl.e0, nloc = dl.nloc;
    float (*s_tpart)[kS2Cons][2][32] = reinterpret_cast<float (*)[kS2Cons][2][32]>(smem + sp.tpart_off);
    const uint32_t xb = sbase + (uint32_t)sp.xr_off;
#define dbg p.dbg   /* re-read from the descriptor in shared memory at every stamp: no register lives across the unit kernels */
    S2_STAMP(dbg, 0);
    XAddr sm;
    {
        const XLayout XL = x_layout(p.K);
        sm.sx = xb + XL.sx;
        sm.x16 = xb + XL.x16;
        sm.zero = sbase + 96u;
    }
    // logical tile T of the phase -> segment
    auto seg_of = [&](int TT) {
        int es = 0;
        if (!swiglu)
            while (es + 1 < p.n_seg && TT >= p.seg[es].n_tiles) { TT -= p.seg[es].n_tiles; es++; }
        return es;
    };
    // Work is dealt in JOBS, one warp each: an entry (32 rows x C chunks) is s_J jobs along K (a two-chunk entry: one job of two
    // chunks or two of one) times s_R jobs along the rows (both 16-row blocks or one), chosen per phase by the host so that the
    // last round of the 14 warps is full.  Job jj = (J R) i + sub of the CTA's range goes to warp jj mod 14; (T, part, ce) locate entry i.
    const int jsh = p.s_jsh, njobs = nloc << jsh;
    int jj = warp, i = jj >> jsh, T = 0, part = 0, ce = 0, s = 0;
    if (jj < njobs) {
        const int e = e0 + i;
        T = e / per_tile;
        const int r = e - T * per_tile;
        part = r / ept;
        ce = r - part * ept;
        s = seg_of(T);
    }
    uint32_t slot, round;
    {
        const uint32_t q = cs.seq0 + (uint32_t)i;
        slot = q % (uint32_t)rg.n_slots;
        round = q / (uint32_t)rg.n_slots;
    }
    int type = -1;
    LaneT lt{0u, 0u, 0u, 0u};
    int ginv = lane & 7;   // lane group that computes row (lane & 7) of a row block (units2.cuh: s2_row_perm)
    uint32_t xtok = 0u;   // becomes an opaque zero after the wait for x: no shared-memory read of x is hoisted above it
    auto load_mat = [&]() {
        const MSeg& msg = p.seg[swiglu ? part : s];
        const int ty = msg.type;
        if (ty != type) {   // (within a phase a type has one pitch; the rows of a tile all come from matrices of one type)
            type = ty;
            const XLayout XL = x_layout(p.K);
            const uint32_t p0 = xb + XL.p0, p1 = xb + XL.p1, p2 = xb + XL.p2;
            lt = (type == T_Q6_K) ? lane_t_q6k(p0, p1, p2, g, t) : (type == T_Q8_0) ? lane_t_q80(p0, p1, p2, g, t) : lane_t_k45(p0, p1, p2, g, t);
            lt.c1 += xtok;
            lt.c2 += xtok;
            lt.grow = s2_row_perm((uint32_t)msg.s_pitch, g);
            ginv = s2_row_perm_inv((uint32_t)msg.s_pitch, lane & 7);
        }
    };
    if (jj < njobs) load_mat();

    // ---- the phase's input: staged by its producer, copied by the loader warp; everything above overlapped the boundary ----
    s_wait(xfull, xpar, s_dead, p.err, 6000, epoch);
    if (p.n_sum > 0 && !p.ll_red) s2_fold_stage(p, smem + sp.xr_off, s_red, xstg);
    S2_STAMP(dbg, 1);
    {
        xtok = smem_token();
        sm.sx += xtok;
        sm.x16 += xtok;
        sm.zero += xtok;
        lt.c1 += xtok;
        lt.c2 += xtok;
    }

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int wrote = 0;   // parts of the current tile this warp has already left in the tile slot
    bool first = true;
#ifdef B200_S2_DEBUG
    unsigned int t_wait = 0;   // SM clocks spent waiting for ring entries
#endif
    while (jj < njobs) {
        {
            const MSeg& wsg = p.seg[swiglu ? part : s];
            const uint32_t RS = (uint32_t)wsg.s_pitch, cbytes = (uint32_t)wsg.chunk_bytes;
            const int sub = jj & ((1 << jsh) - 1), rsh = p.s_R == 2 ? 1 : 0;
#if B200_S2_CONS == 8
            const int rt0 = 0, rt1 = 2;   // (whole units only: the row-block loop of the unit kernels is unrolled)
#else
            const int rt0 = rsh ? (sub & 1) : 0, rt1 = rsh ? rt0 + 1 : 2;
#endif
            const int sC = p.s_C, cpj = sC >> (jsh - rsh);           // chunks per entry / per job
            const int cj = (sub >> rsh) * cpj, c0 = ce * sC;          // first chunk of the job within the entry; of the entry within the row
            const uint32_t e00 = (uint32_t)(c0 + cj) * kMmaChunk;
            const uint32_t doff = ((uint32_t)c0 * cbytes) & 15u;     // the box starts 16-byte aligned (Q6_K: any even residue)
            // A slot always serves the same warps within a phase (n_slots = 14, (J R) x 14 = 0 mod 14) and phases are separated by
            // the grid barrier: "parity r of `full`" can never be a stale phase (stream.cuh needed a guard wait on `empty` here)
#ifdef B200_S2_DEBUG
            t_wait -= (unsigned int)clock();
#endif
            s_wait(rg.full + 8u * slot, round & 1u, s_dead, p.err, 2000 + warp, slot);
#ifdef B200_S2_DEBUG
            t_wait += (unsigned int)clock();
#endif
            if (first) { S2_STAMP(dbg, 2); first = false; }
            const uint32_t a0 = rg.base + slot * (uint32_t)kS2SlotBytes + doff + (uint32_t)cj * cbytes + smem_token();
            for (int c = 0; c < cpj; c++) {
                const uint32_t a = a0 + (uint32_t)c * cbytes, ee = e00 + (uint32_t)c * kMmaChunk;
                switch (type) {
                    case T_Q4_K: S2_UNIT(k45<false>)(a, RS, ee, sm, lt, g, t, rt0, rt1, acc); break;
                    case T_Q5_K: S2_UNIT(k45<true>)(a, RS, ee, sm, lt, g, t, rt0, rt1, acc); break;
                    case T_Q6_K:
                        if (((a - rg.base) & 3u) == 0u) S2_UNIT(q6k<4>)(a, RS, ee, sm, lt, g, t, rt0, rt1, acc);   // (slots and 16 x pitch are 16-byte multiples)
                        else S2_UNIT(q6k<2>)(a, RS, ee, sm, lt, g, t, rt0, rt1, acc);
                        break;
                    default: S2_UNIT(q80)(a, RS, ee, min(wsg.cb, wsg.nb_row - (int)(ee >> 5)), sm, lt, g, t, rt0, rt1, acc); break;
                }
            }
            pin4(acc);   // every shared-memory read of the job has completed before the slot is handed back
            __syncwarp();
            if (lane == 0)   // a slot is free after 4 arrivals: one job arrives 4 / (jobs per entry) times
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(rg.empty + 8u * slot), "r"(4 >> jsh) : "memory");
        }

        // ---- advance the cursor by 14 jobs ----
        const int cur_T = T, cur_part = part;
        jj += kS2Cons;
        {
            const int ni = jj >> jsh, di = ni - i;
            i = ni;
            ce += di;
            slot += (uint32_t)di;
            if (slot >= (uint32_t)rg.n_slots) { slot -= (uint32_t)rg.n_slots; round++; }
        }
        while (ce >= ept) {
            ce -= ept;
            if (++part == p.s_parts) { part = 0; T++; }
        }
        const bool more = jj < njobs;
        if (!more) S2_STAMP(dbg, 3);
        const bool tile_end = !more || T != cur_T;
        if (tile_end || part != cur_part) {
            // ---- this warp leaves (tile cur_T, part cur_part): its 32 row sums go to the tile's slot ----
            const uint32_t ts = cs.tseq0 + (uint32_t)(cur_T - dl.T_first);
            const int tslot = (int)(ts & (kS2TileSlots - 1));
            if (wrote == 0) {   // first store into the slot for this tile: its previous tile (4 tiles ago) must have been merged
                const unsigned int gen = ts >> 2;
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
            s_tpart[tslot][warp][cur_part][lane] = rows32_from_acc(acc, lane, ginv);
#pragma unroll
            for (int k = 0; k < 4; k++) acc[k] = 0.f;
            wrote |= 1 << cur_part;
            if (tile_end) {
                if (swiglu && wrote != 3) s_tpart[tslot][warp][wrote == 1 ? 1 : 0][lane] = 0.f;   // this warp had no job of the other matrix
                wrote = 0;
                __syncwarp();
                if (lane == 0) {   // the loader warp adds the pieces (in job order) when the tile's count is complete
                    __threadfence_block();
                    atomicAdd(&s_tcnt[tslot], 1);
                }
            }
            if (more) {   // (after the flush: the row permutation of the finished tile's matrix was still needed)
                if (T != cur_T) s = seg_of(T);
                load_mat();
            }
        }
    }
    S2_STAMP(dbg, 4);
#ifdef B200_S2_DEBUG
    if (dbg && lane == 0) dbg[((size_t)blockIdx.x * 16 + warp) * 8 + 5] = (unsigned long long)t_wait;
#endif
#undef dbg
    cs.seq0 += (uint32_t)nloc;
    cs.tseq0 += (uint32_t)dl.n_ltiles;
}

// ---------------------------------------------------------------- loader warp, GEMV phase: merge, cross-CTA packets, epilogues
// Runs while the consumers compute: for every tile of this CTA, in order, wait until all contributing warps have left their
// row sums, add them in entry order (fixed: results are run-to-run identical), then
//   * a piece whose head lives in an earlier CTA is published as 32 (value, epoch) packets (the owner polls them);
//   * a tile whose tail lives in later CTAs collects their packets (they were computed FIRST by those CTAs);
//   * the epilogue: 1/rms, bias, residual, SwiGLU, store, staged int8 form of the output, argmax candidates (vocab head).
// The operands of a tile's epilogue (bias / residual / norm weight of the output) are requested BEFORE the tile's count is
// waited for.  Residuals are at least two phases old, so none of this depends on the boundary just crossed.
struct S2Best {
    float v;
    int i;
};
__device__ __forceinline__ int s2_gemv_epilogues(const MParams& p, const Stream2Params& sp, uint8_t* smem, uint32_t xfull, uint32_t xpar,
                                                  uint32_t xstg, const float* s_red, uint32_t& n_fold, uint32_t tseq0, int* s_tcnt, volatile unsigned int* s_tdone,
                                                  volatile int* s_dead, unsigned int epoch, bool greedy, S2Best& best, unsigned int ll_ep) {
    const int lane = threadIdx.x & 31;
    const bool swiglu = p.epi == ME_SWIGLU;
    const int per_tile = p.s_parts * p.s_ept;
    const S2Deal dl = s2_deal(p);
    const int e0 = dl.e0, e1 = dl.e0 + dl.nloc;
    const bool fold = p.n_sum > 0 && !p.ll_red;   // (tensor parallel) the consumers stage the input themselves: see s2_gemv_cta
    const uint32_t fold_par = n_fold & 1u;
    if (fold) n_fold++;
    if (dl.n_ltiles == 0) {
        if (greedy && p.cand && lane == 0) {   // no tile of the vocab head here: no candidate
            sp.cand_val[blockIdx.x] = -INFINITY;
            sp.cand_idx[blockIdx.x] = -1;
        }
        return 0;
    }
    float (*s_tpart)[kS2Cons][2][32] = reinterpret_cast<float (*)[kS2Cons][2][32]>(smem + sp.tpart_off);
    unsigned long long* const dbg = p.dbg;
    float unscale = 1.0f;
    if (p.norm_w) {   // sum of x^2 from the per-group partial sums of the staged input: the loader waits for its own copy
        float tot = 0.0f;
        if (fold) {
            s_wait(xstg, fold_par, s_dead, p.err, 6450, epoch);
            if (lane < kS2Cons) tot = *(volatile const float*)(s_red + lane);
        } else {
            s_wait(xfull, xpar, s_dead, p.err, 6400, epoch);
            const uint32_t ssq = smem_u32(smem) + (uint32_t)sp.xr_off + x_layout(p.K).ssq + smem_token();
            for (int k = lane; k < (p.K >> 5); k += 32) tot += lds_f32(ssq + 4u * (uint32_t)k);
        }
        tot = warp_sum(tot);
        unscale = 1.0f / sqrtf(tot / (float)p.K + p.eps);
    }
    const bool cand = greedy && p.cand;
    for (int tl = 0; tl < dl.n_ltiles; tl++) {
        const int T = dl.T_first + tl;
        int es = 0, etile = T;
        if (!swiglu)
            while (es + 1 < p.n_seg && etile >= p.seg[es].n_tiles) { etile -= p.seg[es].n_tiles; es++; }
        const MSeg& sg = p.seg[es];
        const int j = etile * kMmaRows + lane;
        const bool valid = j < sg.n_rows;
        const bool head_local = T * per_tile >= e0, tail_local = (T + 1) * per_tile <= e1;
        float e_bias = 0.0f, e_res = 0.0f, e_w = 1.0f;
        if (head_local) {   // this CTA finishes the tile: its epilogue operands, in flight while the tile is computed
            if (valid && sg.bias) e_bias = sg.bias[j];
            if (valid && p.epi == ME_RESIDUAL) e_res = __ldcg(p.residual + j);
            if (p.stage_out && p.stage_w && !p.ll_red && j < p.stage_K) e_w = p.stage_w[j];
        }
        const int t_lo = (max(T * per_tile, e0) - e0) << p.s_jsh, t_hi = (min((T + 1) * per_tile, e1) - e0) << p.s_jsh;   // local jobs of the tile
        const int n_cw = min(t_hi - t_lo, kS2Cons);   // warps that hold a piece of it
        const uint32_t ts = tseq0 + (uint32_t)tl;
        const int tslot = (int)(ts & (kS2TileSlots - 1));
        {
            const long long w0 = clock64();
            while (*(volatile int*)&s_tcnt[tslot] != n_cw) {
                if (*s_dead) break;
                if (clock64() - w0 > 2000000000LL) {
                    *s_dead = 1;
                    if (atomicExch(p.err, 7) == 0) { p.err[1] = 7500; p.err[2] = (int)blockIdx.x; p.err[3] = (int)ts; }
                    break;
                }
            }
        }
        __threadfence_block();
        float vg = 0.f, vu = 0.f;
        {   // all loads first (independent), then the additions in job order: fixed, whoever arrived last
            float pg[kS2Cons], pu[kS2Cons];
            int w = t_lo % kS2Cons;
#pragma unroll
            for (int k = 0; k < kS2Cons; k++) {
                pg[k] = (k < n_cw) ? s_tpart[tslot][w][0][lane] : 0.f;
                pu[k] = (k < n_cw && swiglu) ? s_tpart[tslot][w][1][lane] : 0.f;
                if (++w == kS2Cons) w = 0;
            }
#pragma unroll
            for (int k = 0; k < kS2Cons; k++) {
                if (k < n_cw) { vg += pg[k]; vu += pu[k]; }
            }
        }
        __syncwarp();
        if (lane == 0) {
            s_tcnt[tslot] = 0;
            __threadfence_block();
            s_tdone[tslot] = (ts >> 2) + 1u;
        }
        if (tl == 0) S2_STAMP(dbg, 4);
        if (!head_local) {   // the tile's head lives in an earlier CTA: publish this piece
            uint2* mine = sp.ll + (size_t)blockIdx.x * 64;
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(mine + lane), "r"(__float_as_uint(vg)), "r"(epoch) : "memory");
            if (swiglu) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(mine + 32 + lane), "r"(__float_as_uint(vu)), "r"(epoch) : "memory");
            continue;
        }
        if (!tail_local) {   // its tail lives in later CTAs
            const long long nb = gridDim.x, E = p.s_E;
            const int h_last = (int)((((long long)(T + 1) * per_tile) * nb - 1) / E);   // CTA of the tile's last entry
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
                        if (atomicExch(p.err, 6) == 0) { p.err[1] = 8000; p.err[2] = (int)blockIdx.x; p.err[3] = h; }
                        break;
                    }
                }
                vg += __uint_as_float(a0);
                vu += __uint_as_float(b0);
            }
        }
        // ---- epilogue: lane L owns row etile * 32 + L of segment es (vu: the up row for SwiGLU) ----
        vg *= unscale;
        float val = vg;
        if (swiglu) val = mma_silu(vg) * (vu * unscale);
        if (valid) {
            val += e_bias;
            val += e_res;
            if (cand) {   // raw-logit argmax, LAST maximal index wins (src/main.rs:1816-1821)
                if (val > best.v || (val == best.v && j > best.i) || best.i < 0) { best.v = val; best.i = j; }
            } else if (p.n_peer > 0 && p.ll_red) {   // partial of a row-parallel GEMV as (value, epoch) packets to every rank (NVLink peer memory)
                for (int r = 0; r < p.n_peer; r++)
                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(p.peer_out[r]) + j), "r"(__float_as_uint(val)), "r"(ll_ep) : "memory");
            } else if (p.n_peer > 0) {   // ... as plain floats; summed by the REDUCE phase after the flag exchange
                for (int r = 0; r < p.n_peer; r++) p.peer_out[r][j] = val;
            } else {
                sg.out[j] = val;
            }
        }
        if (p.stage_out && !p.ll_red) stage_out32(val, e_w, j, p.stage_K, p.stage_out);
    }
    S2_STAMP(dbg, 5);
    if (cand) {   // the CTA's candidate of this token: larger value, then larger index
        float bv = best.v;
        int bi = best.i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            sp.cand_val[blockIdx.x] = bv;
            sp.cand_idx[blockIdx.x] = bi;
        }
        best.v = -INFINITY;
        best.i = -1;
    }
    return dl.n_ltiles;
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
#ifndef B200_ATTN2_UB
#define B200_ATTN2_UB 2
#endif
    constexpr int UB = GMAX <= 4 ? B200_ATTN2_UB : 1;   // positions per batch (two batches in flight per warp); lab: -DB200_ATTN2_UB=4
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
    float kA[UB][VEC], vA[UB][VEC], kB[UB][VEC], vB[UB][VEC];
    int pos0 = start + warp;
    if (active && pos0 < end_g) load(pos0, kA, vA);
#ifndef B200_NO_ATTN_PRE2
    if (active && pos0 + STEP < end_g) load(pos0 + STEP, kB, vB);   // (32 positions per split = two batches per warp: nothing is left to fetch after the boundary)
#endif

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
#ifdef B200_NO_ATTN_PRE2
    if (pos0 + STEP < end_g) load(pos0 + STEP, kB, vB);
#endif
    while (pos0 < end_g) {   // kA = batch at pos0, kB = batch at pos0 + STEP (both in flight or landed)
        compute(pos0, end_g, kA, vA);
        if (pos0 + 2 * STEP < end_g) load(pos0 + 2 * STEP, kA, vA);
        pos0 += STEP;
        if (pos0 >= end_g) break;
        compute(pos0, end_g, kB, vB);
        if (pos0 + 2 * STEP < end_g) load(pos0 + 2 * STEP, kB, vB);
        pos0 += STEP;
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
// tp_ep: (tensor parallel) number of this cross-GPU exchange within the launch, the same on every rank
__device__ __forceinline__ void s2_pick(const Stream2Params& sp, float* s_av, int* s_ai, int* s_tok, unsigned int tp_ep = 0) {
    const MegaParams& mp = sp.mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)gridDim.x;   // one candidate per CTA (its loader warp ran the vocab head's epilogues)
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
        if (mp.tp_size > 1) {   // every rank picks the same winner among the per-rank candidates (ties: largest global index)
            const unsigned int ep = mp.tp_epoch0 + tp_ep;
            if (bi >= 0) bi += mp.tp_rank * mp.vocab_local;
            for (int r = 0; r < mp.tp_size; r++) {
                float* dst = mp.tp_peer_cand[r] + 2 * mp.tp_rank;
                asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(best) : "memory");
                asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + 1), "f"(__int_as_float(bi)) : "memory");
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int r = 0; r < mp.tp_size; r++)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mp.tp_peer_flags[r] + mp.tp_rank), "r"(ep) : "memory");
            const long long t0 = clock64();
            best = -INFINITY;
            bi = -1;
            for (int r = 0; r < mp.tp_size; r++) {
                for (;;) {
                    unsigned int fv;
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(fv) : "l"(mp.tp_flags + r) : "memory");
                    if ((int)(fv - ep) >= 0) break;
                    if (clock64() - t0 > 3000000000LL) { atomicExch(mp.err, 3); break; }
                }
                float cv, ci;
                asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(cv) : "l"(mp.tp_cand + 2 * r) : "memory");
                asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(ci) : "l"(mp.tp_cand + 2 * r + 1) : "memory");
                const int ii = __float_as_int(ci);
                if (ii >= 0 && (bi < 0 || cv > best || (cv == best && ii > bi))) { best = cv; bi = ii; }
            }
        }
        mp.st->token = bi;
        const int gcount = mp.st->n_generated;
        if (gcount < mp.max_generated) mp.generated[gcount] = bi;
        mp.st->n_generated = gcount + 1;
        *s_tok = bi;
    }
    s2_cons_sync();
}

__device__ __forceinline__ void s2_embed(const MParams& p, const Stream2Params& sp, bool pick, float* s_av, int* s_ai, int* s_tok, unsigned int tp_ep) {
    const MegaParams& mp = sp.mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (pick) {
        s2_pick(sp, s_av, s_ai, s_tok, tp_ep);
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

// ---------------------------------------------------------------- REDUCE phase (tensor parallel)
// The all-reduce of a row-parallel GEMV, finished locally and identically on every rank: x[j] = sum over ranks (rank order) of the partial
// vectors the ranks left in this GPU's buffer + the residual; stored as f32 (it is the residual of a later phase) and in the staged form the
// next GEMV's bulk copy reads.  32 consecutive elements per warp step (the staged form needs whole groups).
__device__ __forceinline__ void s2_reduce_phase(const MParams& p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_grp = p.K >> 5;
    for (int grp = (int)blockIdx.x * kS2Cons + warp; grp < n_grp; grp += (int)gridDim.x * kS2Cons) {
        const int j = grp * 32 + lane;
        float v = __ldcg(p.xsum + j);
        for (int r = 1; r < p.n_sum; r++) v += __ldcg(p.xsum + (size_t)r * p.sum_stride + j);
        if (p.x_res) v += __ldcg(p.x_res + j);
        if (p.x_full_out) p.x_full_out[j] = v;
        stage_out32(v, p.stage_w ? p.stage_w[j] : 1.0f, j, p.stage_K, p.stage_out);
    }
}

// ---------------------------------------------------------------- all-reduce inside a row-parallel GEMV phase (tensor parallel, ll_red)
// Consumer warps, after their last job: one group of 32 elements per warp and step.  The P partial values of an element arrive as 8-byte
// (value, epoch) packets written by the loader warps of every rank (this one included) straight from their epilogues: a packet is valid
// when its epoch is this exchange's -- no fence, no flag round trip, no grid barrier in front of the reduction.  Sum in rank order +
// residual (identical on every rank), f32 vector (a later residual) and the staged form the next GEMV's bulk copy reads.
__device__ __noinline__ void s2_reduce_ll(const MParams& p, unsigned int ep, volatile int* s_dead) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_grp = p.stage_K >> 5;
    const uint2* base = reinterpret_cast<const uint2*>(p.xsum);
    for (int grp = (int)blockIdx.x * kS2Cons + warp; grp < n_grp; grp += (int)gridDim.x * kS2Cons) {
        const int j = grp * 32 + lane;
        const float res = p.x_res ? __ldcg(p.x_res + j) : 0.0f;
        const float sw = p.stage_w ? p.stage_w[j] : 1.0f;
        uint32_t val[kMmaMaxPeers];
        const long long t0 = clock64();
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int r = 0; r < kMmaMaxPeers; r++) {
                if (r < p.n_sum) {
                    uint32_t e;
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(val[r]), "=r"(e) : "l"(base + (size_t)r * p.sum_stride + j) : "memory");
                    ok = ok && e == ep;
                }
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (*s_dead || clock64() - t0 > 3000000000LL) {
                *s_dead = 1;
                if (atomicExch(p.err, 3) == 0) { p.err[1] = 5100; p.err[2] = (int)blockIdx.x; p.err[3] = (int)ep; }
                break;
            }
        }
        float v = __uint_as_float(val[0]);
#pragma unroll
        for (int r = 1; r < kMmaMaxPeers; r++)
            if (r < p.n_sum) v += __uint_as_float(val[r]);
        v += res;
        if (p.x_full_out) p.x_full_out[j] = v;
        stage_out32(v, sw, j, p.stage_K, p.stage_out);
    }
}

// ---------------------------------------------------------------- loader warp: the phase boundary, then the phase's epilogues
__device__ __forceinline__ void s2_loader(const Stream2Params& sp, uint8_t* smem, uint32_t xfull, uint32_t done, int* s_pos, float* s_rope,
                                          int hd, int* s_tcnt, volatile unsigned int* s_tdone, volatile int* s_dead, uint32_t xstg, const float* s_red) {
    const MegaParams& mp = sp.mp;
    const int lane = threadIdx.x & 31;
    uint32_t n_fold = 0;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const long long total = (long long)mp.n_tokens * n_run;
    const bool greedy = mp.mode == MEGA_GREEDY;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t xr = sbase + (uint32_t)sp.xr_off;
    const MegaPhase* s_desc = reinterpret_cast<const MegaPhase*>(smem + sp.desc_off);
    unsigned int target = 0;
    uint32_t tseq0 = 0;
    S2Best best{-INFINITY, -1};
    int kv_len = 1;
    int ph = 0;
    bool prev_tp = false;          // the phase that just ended left partial sums in the peers' memory
    unsigned int tp_tok = 0, tp_k = 0;   // token of the launch, row-parallel exchanges of that token so far
    for (long long gb = 0; gb < total + (greedy ? 1 : 0); gb++) {
        const bool last = gb == total;     // the boundary after the last vocab head (greedy: CTA 0 picks the last token)
        const MegaPhase& cur = s_desc[gb & 1];
        unsigned long long* ldbg = nullptr;   // stamps of the boundary BEFORE a phase go to that phase's debug buffer
        __syncwarp();   // the epilogue stores of every lane are ordered before lane 0's release below
        if (lane == 0) {
            uint32_t tx = 0;
            if (gb > 0) {
                s_wait(done, (uint32_t)((gb - 1) & 1), s_dead, mp.err, 9000, (uint32_t)gb);
                s_wait(xfull, (uint32_t)((gb - 1) & 1), s_dead, mp.err, 9100, (uint32_t)gb);   // s_desc[gb & 1] has landed
            }
            if (!last && cur.kind == PH_GEMV && cur.gemv.dbg) ldbg = cur.gemv.dbg + ((size_t)blockIdx.x * 16 + kS2LoaderWarp) * 8;
            if (ldbg) ldbg[0] = (unsigned long long)clock64();
            if (!last && gb + 1 < total) {   // descriptor of the NEXT phase: constant data, issued before the barrier is polled
                int nph = ph + 1;
                if (nph == n_run) nph = 0;
                tx += (uint32_t)sizeof(MegaPhase);
                bulk_g2s(sbase + (uint32_t)sp.desc_off + (uint32_t)(((gb + 1) & 1) * sizeof(MegaPhase)), mp.phases + nph,
                         (uint32_t)sizeof(MegaPhase), xfull);
            }
            // what the phase's input is: known before the barrier opens, so the mbarrier is armed first and the copy is the only
            // thing left to do afterwards.  Phases without an input copy (EMBED, idle CTAs of an attention phase, the final pick)
            // must not open before the grid barrier has: their arrive comes after the poll.
            const void* src[3] = {nullptr, nullptr, nullptr};
            uint32_t dst[3] = {xr, 0u, 0u}, nby[3] = {0u, 0u, 0u};
            if (!last && cur.kind == PH_GEMV) {
                src[0] = cur.gemv.x_staged;
                nby[0] = (uint32_t)cur.gemv.x_bytes;
            } else if (!last && cur.kind == PH_ATTN && ph != 1) {
                const AttnParams& ap = cur.attn;
                const int ns = attn_eff_splits(kv_len, ap.n_splits, ap.min_chunk);
                if ((int)blockIdx.x < ap.n_kv * ns) {
                    const int kh = (int)blockIdx.x / ns, gmax = ap.G <= 4 ? 4 : 8;
                    src[0] = ap.qkv_raw + (size_t)kh * ap.G * hd;                                    nby[0] = (uint32_t)(ap.G * hd * 4);
                    src[1] = ap.qkv_raw + (size_t)ap.n_heads * hd + (size_t)kh * hd;                nby[1] = (uint32_t)(hd * 4);
                    src[2] = ap.qkv_raw + (size_t)(ap.n_heads + ap.n_kv) * hd + (size_t)kh * hd;    nby[2] = (uint32_t)(hd * 4);
                    dst[1] = xr + (uint32_t)(gmax * hd * 4);
                    dst[2] = xr + (uint32_t)((gmax + 1) * hd * 4);
                }
            }
            const bool early = nby[0] != 0u;
            tx += nby[0] + nby[1] + nby[2];
            if (early) mbar_arrive_expect_tx(xfull, tx);
            if (gb > 0) {   // grid barrier: everything every CTA wrote in phase gb - 1 is visible after this
                target += gridDim.x;
                if (prev_tp) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(mp.bar) : "memory");   // stores to peer memory included
                else asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(mp.bar) : "memory");
                if (ldbg) ldbg[1] = (unsigned long long)clock64();
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
                if (prev_tp) {
                    // every CTA of THIS GPU has arrived, i.e. this rank's partial sums are in every peer's memory: CTA 0 says so to
                    // the peers; every CTA waits (in local memory) until every rank has said so
                    const unsigned int ep = mp.tp_epoch0 + tp_tok * (unsigned int)sp.tp_per_token + (++tp_k);
                    if (blockIdx.x == 0)
                        for (int r = 0; r < mp.tp_size; r++)
                            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mp.tp_peer_flags[r] + mp.tp_rank), "r"(ep) : "memory");
                    for (int r = 0; r < mp.tp_size; r++) {
                        for (;;) {
                            unsigned int fv;
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(fv) : "l"(mp.tp_flags + r) : "memory");
                            if ((int)(fv - ep) >= 0) break;
                            if (*s_dead || clock64() - t0 > 3000000000LL) {
                                *s_dead = 1;
                                if (atomicExch(mp.err, 3) == 0) { mp.err[1] = 5000 + r; mp.err[2] = (int)blockIdx.x; mp.err[3] = (int)ep; }
                                break;
                            }
                        }
                    }
                }
                if (mp.dbg && blockIdx.x == 0) mp.dbg[last ? n_run : ph] = gtimer();   // boundary BEFORE phase ph of the current token
                if (ldbg) ldbg[2] = (unsigned long long)clock64();
                asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes of other CTAs -> this thread's bulk copies
            }
            if (!last && ph == 1) {   // first phase after the embedding: the token's position (written by CTA 0 in the EMBED phase)
                const int pc = __ldcg(&mp.st->pos_cur);
                *s_pos = pc;
                __threadfence_block();   // (the consumers' acquire is the copy's completion, not this thread's arrive)
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (nby[k]) bulk_g2s(dst[k], src[k], nby[k], xfull);
            if (!early) mbar_arrive_expect_tx(xfull, tx);
            if (ldbg) ldbg[3] = (unsigned long long)clock64();
        }
        __syncwarp();
        if (last) break;
        if (ph == 1) {   // the rotation angles of this token (Backend::rope, cpu/ops.rs:1216-1337): needed from phase 2 on
            kv_len = *(volatile int*)s_pos + 1;
            const AttnParams& ap = reinterpret_cast<const MegaPhase*>(mp.phases + 2)->attn;
            const float position = (float)(kv_len - 1) / ap.rope_scale;
            for (int pi = lane; pi < hd / 2; pi += 32) {
                const float theta = position * ap.freq[pi];
                s_rope[pi] = cosf(theta);
                s_rope[hd / 2 + pi] = sinf(theta);
            }
            __syncwarp();
        }
        prev_tp = cur.tp_sync != 0;
        if (cur.kind == PH_GEMV) {
            unsigned int ll_ep = 0u;   // (ll_red) the exchange's epoch: the same count on every rank, in every loader and consumer warp
            if (cur.gemv.ll_red) ll_ep = mp.tp_epoch0 + tp_tok * (unsigned int)sp.tp_per_token + (++tp_k);
            tseq0 += (uint32_t)s2_gemv_epilogues(cur.gemv, sp, smem, xfull, (uint32_t)(gb & 1), xstg, s_red, n_fold, tseq0, s_tcnt, s_tdone, s_dead,
                                                 sp.epoch0 + (unsigned int)gb + 1u, greedy, best, ll_ep);
        }
        if (++ph == n_run) { ph = 0; tp_tok++; tp_k = 0; }
    }
}

// ---------------------------------------------------------------- the kernel
template <int HD, int GMAX>
__global__ void __launch_bounds__(kS2Threads, 1) stream2_decode_kernel(const __grid_constant__ Stream2Params sp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bars[2 * kS2MaxSlots + 3];
    __shared__ int s_tcnt[kS2TileSlots];
    __shared__ unsigned int s_tdone[kS2TileSlots];
    __shared__ unsigned int s_ticket;
    __shared__ int s_dead;
    __shared__ int s_pos;
    __shared__ int s_tok;
    __shared__ float s_rope[HD];
    __shared__ float s_av[kS2Cons];
    __shared__ int s_ai[kS2Cons];
    __shared__ float s_red[kS2Cons];

    const MegaParams& mp = sp.mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SRing rg;
    rg.base = smem_u32(smem) + (uint32_t)sp.ring_off;
    rg.full = smem_u32(s_bars);
    rg.empty = rg.full + 8u * (uint32_t)sp.n_slots;
    rg.n_slots = sp.n_slots;
    const uint32_t xfull = rg.full + 16u * (uint32_t)kS2MaxSlots, done = xfull + 8u, xstg = done + 8u;
    if (tid == 0) {
        for (int i = 0; i < sp.n_slots; i++) {
            mbar_init(rg.full + 8u * i, 1);
            mbar_init(rg.empty + 8u * i, 4);   // both warps of the pair(s) that computed the entry's chunks hand the slot back (see s2_gemv_cta)
        }
        mbar_init(xfull, 1);
        mbar_init(done, kS2Cons);
        mbar_init(xstg, 1);
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

    // 16 warps at 128 registers cannot hold a unit kernel plus the loop state without spilling (and the carve-out leaves little
    // L1 to spill into): the service warpgroup hands registers to the three consumer warpgroups
    // (the two register budgets must be separate regions of the program: code after a join would be compiled for the smaller one)
    if (warp >= kS2Cons) {
        if (kS2Regs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kS2ServRegs));
        if (warp >= kS2ProdWarp && warp < kS2ProdWarp + kS2Prods) {
            if (lane == 0) s2_producer(sp, rg, &s_dead, warp - kS2ProdWarp);
            s_drain(&s_dead);
        } else if (warp == kS2LoaderWarp) {
            s2_loader(sp, smem, xfull, done, &s_pos, s_rope, HD, s_tcnt, s_tdone, &s_dead, xstg, s_red);
            s_drain(&s_dead);
        }
        return;
    }
    if (kS2Regs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kS2ConsRegs));

    // ---------------- consumer warps ----------------
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const bool greedy = mp.mode == MEGA_GREEDY;
    const MegaPhase* s_desc = reinterpret_cast<const MegaPhase*>(smem + sp.desc_off);
    S2Cons cs{0u, 0u};
    long long gph = 0;
    for (int tok = 0; tok < mp.n_tokens; tok++) {
        unsigned int tp_k = 0;   // (ll_red) row-parallel exchanges of this token so far
        for (int ph = 0; ph < n_run; ph++, gph++) {
            const MegaPhase& cur = s_desc[gph & 1];
            const uint32_t xpar = (uint32_t)(gph & 1);
            const unsigned int epoch = sp.epoch0 + (unsigned int)gph + 1u;
            if (cur.kind == PH_GEMV) {
                s2_gemv_cta(cur.gemv, sp, smem, rg, cs, xfull, xpar, xstg, s_red, s_tcnt, s_tdone, &s_dead, epoch);
                if (cur.gemv.ll_red) s2_reduce_ll(cur.gemv, mp.tp_epoch0 + (unsigned int)tok * (unsigned int)sp.tp_per_token + (++tp_k), &s_dead);
            } else if (cur.kind == PH_ATTN) {
                attn2_phase<HD, GMAX, kS2Cons>(cur.attn, s_pos + 1, reinterpret_cast<float*>(smem + sp.xr_off), xfull, xpar, &s_dead, mp.err,
                                               epoch, &s_ticket, s_rope);
            } else if (cur.kind == PH_REDUCE) {
                s_wait(xfull, xpar, &s_dead, mp.err, 6250, epoch);
                s2_reduce_phase(cur.gemv);
            } else {
                s_wait(xfull, xpar, &s_dead, mp.err, 6200, epoch);
                if (blockIdx.x == 0) s2_embed(cur.gemv, sp, greedy && tok > 0, s_av, s_ai, &s_tok, (unsigned int)(tok * sp.tp_per_token));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(done);
        }
    }
    if (greedy) {   // the last token's pick, after the boundary that follows its vocab head
        s_wait(xfull, (uint32_t)(gph & 1), &s_dead, mp.err, 6300, (uint32_t)gph);
        if (blockIdx.x == 0) s2_pick(sp, s_av, s_ai, &s_tok, (unsigned int)(mp.n_tokens * sp.tp_per_token));
    }
    s_drain(&s_dead);
}

}  // namespace b200
