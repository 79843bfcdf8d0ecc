// stream.cuh — the streamed per-token megakernel: one persistent CTA per SM whose PRODUCER warp never stops pulling
// weights from HBM while the consumer warps walk the phases of the token.
//
// Why (profiles/r01_mega_ncu_full.md, per-phase timeline): in the first megakernel (mega.cuh) every GEMV phase
// pays ~5 us of fixed cost (grid barrier, x staging, ring ramp-up, stragglers) against 1.5-10 us of streaming, and
// HBM idles during all of it: 29 % of the copy roofline for a whole token.  Weights never depend on the token, so
// here they are fetched by a warp that is not part of any phase:
//
//   * producer = one elected lane of warp 8.  It walks the same phase program as the consumers, but only the GEMV
//     phases, and issues ONE `cp.async.bulk.tensor.2d` (TMA, SASS UTMALDG) per ring entry: a box of 32 weight rows x
//     4 super-blocks (Q4_K: 32 x 576 B) straight from the untouched GGUF row layout, described by a tensor map per
//     weight matrix ([n_rows][row_bytes], encoded once at model load).  Rows / columns past the end of a matrix are
//     zero-filled by the TMA unit (a zero block has d = 0: it contributes nothing), so ragged shapes need no code.
//     The only thing the producer ever waits for is a free slot of the ring (mbarrier `empty`), never a grid barrier:
//     while the consumers sit in a barrier, stage x or run attention, the next GEMV's weights keep arriving
//     (6 slots x 148 SMs = up to 24 MB in flight / buffered).
//   * consumers = 8 warps.  A phase's 32-row tiles are dealt whole to CTAs (no tile is shared between CTAs: nothing
//     is merged through global memory, no tickets), the entries of a CTA's tiles are dealt in contiguous runs to its
//     warps; entry n of the CTA's sequence goes to warp n % 8 so the ring is consumed in issue order.  A warp waits
//     for its entry (mbarrier `full`, completed by the TMA transaction bytes), runs the integer tensor-pipe unit
//     kernels of gemv_mma.cuh on it (the box pitch is the ring row stride: 576 = 64 mod 128 keeps the fragment loads
//     conflict-free), and hands the slot back.  Tiles cut across warps are merged in shared memory in warp order
//     (deterministic), epilogues (bias, residual, SwiGLU, staged int8 form of the output for the next GEMV) as before.
//   * attention / RoPE / KV-write phases are attn_decode_item (attention.cuh) on the consumer warps; all CTA-wide
//     synchronisation of the consumers uses named barrier 1 (the producer never joins a barrier).
//
// Replaces the same reference code as mega.cuh: GpuOnlyInference::forward (src/backend/cuda/gpu_only.rs:849-1010),
// CPU LlamaModel::forward (src/model/llama.rs:275-362) with the fused dots of src/backend/cpu/simd.rs:931-1146.
// Dense single-GPU models whose weight rows are 16-byte multiples take this path; the rest stay on mega.cuh.
#pragma once
#include "mega.cuh"

namespace b200 {

constexpr int kSW = 7;                           // consumer warps (7 + the producer warp = 256 threads: 255 registers each, no spills)
constexpr int kSNT = kSW * 32;                  // consumer threads
constexpr int kStreamThreads = kSNT + 32;       // + the producer warp
constexpr int kStreamSlotBytes = 9216;          // 32 rows x 288 bytes (Q4_K, 2 super-blocks): small entries, many slots
constexpr int kStreamMaxSlots = 24;

struct StreamParams {
    MegaParams mp;
    int ring_off;   // bytes of dynamic shared memory before the ring (x staging / attention scratch), multiple of 128
    int n_slots;
    int no_load;    // experiment: the producer arms the barriers but moves no bytes (consumer speed without HBM)
};

// ---------------------------------------------------------------- host-side geometry
// Why small entries: a slot is busy from the moment its TMA load is issued until the consuming warp has finished the
// entry (~1.5-2 us of flight + the warp's compute time), so the ring needs (bytes in flight: ~44 GB/s per SM x 2 us)
// + (warps x entry) bytes.  With 32 x 576-byte entries (first version) 6 slots fed 8 warps at 3.3 TB/s; halving the
// entry halves both the compute hold time and the bytes pinned by computing warps.
// Chunks (256 elements) per ring entry a type would like; a phase uses the minimum over its matrices.
inline int stream_pref_chunks(int type) { return type == T_Q4_K ? 2 : 1; }
// Inner box of the tensor map = row pitch of a ring entry.  The box must start on a 16-byte boundary of the row
// (tools/tma_lab.cu: an unaligned start is an illegal instruction), so an entry that starts at byte b of its row is
// fetched from b & ~15 and the consumer skips b & 15 bytes: the pitch covers the largest such residue.
// Q4_K 288 (C = 2) / 144, Q5_K 176, Q6_K 210 + 14 = 224, Q8_0 272.
inline int stream_pitch(int type, int C) {
    const int cb = kMmaChunk / type_block_elems(type), bb = type_block_bytes(type);
    const int raw = C * cb * bb;
    int maxres = 0;
    for (int ce = 0; ce < 16; ce++) maxres = std::max(maxres, (ce * raw) & 15);
    return (raw + maxres + 15) & ~15;
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSNT) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
// Bounded wait.  A CTA that gives up raises s_dead so that every other wait of the CTA returns at once: all threads
// keep walking the program (barrier counts stay consistent, nothing hangs), the results are garbage and *err says so.
// err[0] = 4, err[1..3] = (which wait, CTA, ring sequence number) of the first wait that gave up (b200_debug_err).
__device__ __forceinline__ bool s_wait(uint32_t bar, uint32_t parity, volatile int* s_dead, int* err, int code, uint32_t seq) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    for (;;) {
        if (mbar_try_wait(bar, parity)) return true;
        if (*s_dead) return false;
        if (clock64() - t0 > 2000000000LL) {  // ~1 s
            *s_dead = 1;
            if (atomicExch(err, 4) == 0) { err[1] = code; err[2] = (int)blockIdx.x; err[3] = (int)seq; }
            return false;
        }
    }
}
// A CTA that gave up may still have TMA loads in flight: let them land before the CTA (and its shared memory) goes away
__device__ __forceinline__ void s_drain(volatile int* s_dead) {
    if (*s_dead) {
        const long long t0 = clock64();
        while (clock64() - t0 < 400000LL) {}
    }
}

// The tiles of a phase this CTA owns, and the entries they make
struct SDeal {
    int T0, cnt, E;
};
__device__ __forceinline__ SDeal s_deal(int s_rot, int s_ncta, int s_cbase, int s_crem, int per_tile) {
    int b = (int)blockIdx.x + s_rot;
    if (b >= (int)gridDim.x) b -= (int)gridDim.x;
    SDeal d{0, 0, 0};
    if (b < s_ncta) {
        d.T0 = b * s_cbase + min(b, s_crem);
        d.cnt = s_cbase + (b < s_crem ? 1 : 0);
    }
    d.E = d.cnt * per_tile;
    return d;
}

struct SRing {
    uint32_t base, full, empty;   // shared-space addresses: slots, full[n_slots], empty[n_slots] (8 bytes each)
    int n_slots;
};

// ---------------------------------------------------------------- producer
// What the producer needs to know about a GEMV phase, loaded from the phase program in global memory one phase AHEAD
// (the loads of phase n + 1 are in flight while the entries of phase n are issued: a descriptor fetch is two dependent
// L2 round trips, ~1.5 us, and would otherwise sit in front of the first copy of every phase).
struct PDesc {
    int gemv;          // 0: not a GEMV phase
    int n_seg, ept, parts, swiglu;
    int rot, ncta, cbase, crem;
    const void* tm[3];
    int nt[3], cstep[3], bytes[3];
};
__device__ __forceinline__ void pdesc_load(PDesc& d, const MegaPhase* P) {
    d.gemv = P->kind == PH_GEMV;
    const MParams* g = &P->gemv;
    d.n_seg = g->n_seg; d.ept = g->s_ept; d.parts = g->s_parts; d.swiglu = g->epi == ME_SWIGLU;
    d.rot = g->s_rot; d.ncta = g->s_ncta; d.cbase = g->s_cbase; d.crem = g->s_crem;
    const int sC = g->s_C;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const MSeg* sg = &g->seg[s];
        d.tm[s] = sg->tmap;
        d.nt[s] = sg->n_tiles;
        d.cstep[s] = sC * sg->chunk_bytes;   // bytes between the entries of a row
        d.bytes[s] = sg->s_pitch * kMmaRows;
    }
}

__device__ __forceinline__ void stream_producer(const StreamParams& sp, const SRing& rg, volatile int* s_dead) {
    const MegaParams& mp = sp.mp;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    const long long total = (long long)mp.n_tokens * n_run;
    uint32_t slot = 0, round = 0;
    PDesc cur, nxt;
    pdesc_load(cur, mp.phases);
    int ph_next = 1 % n_run;
    for (long long it = 0; it < total; it++) {
        pdesc_load(nxt, mp.phases + ph_next);   // issued now, first used after this phase's entries
        if (++ph_next == n_run) ph_next = 0;
        if (cur.gemv) {
            const int ept = cur.ept, per_tile = cur.parts * ept;
            const SDeal d = s_deal(cur.rot, cur.ncta, cur.cbase, cur.crem, per_tile);
            const int base = d.E / kSW, rem = d.E - base * kSW;
            if (it < n_run && d.E > 0) {
                // tensor maps live in global memory (written by the host before the launch): the TMA unit reads them
                // through the tensormap proxy, which needs an acquire fence in every CTA before the first use
#pragma unroll
                for (int s = 0; s < 3; s++)
                    if (s < cur.n_seg) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(cur.tm[s]) : "memory");
            }
            // one cursor per consumer warp (its run of entries is contiguous): tile within the phase, part, entry in the row
            int c_tile[kSW], c_part[kSW], c_ce[kSW];
#pragma unroll
            for (int w = 0; w < kSW; w++) {
                const int j = w * base + min(w, rem);
                const int tl = j / per_tile, r = j - tl * per_tile;
                c_tile[w] = d.T0 + tl;
                c_part[w] = r / ept;
                c_ce[w] = r - c_part[w] * ept;
            }
            for (int i = 0; i <= base; i++) {
                const int nw = (i < base) ? kSW : rem;
#pragma unroll
                for (int w = 0; w < kSW; w++) {
                    if (w < nw) {
                        const int part = c_part[w], ce = c_ce[w];
                        int s = 0, tile = c_tile[w];
                        if (cur.swiglu) {
                            s = part;
                        } else {
                            if (cur.n_seg > 1 && tile >= cur.nt[0]) { tile -= cur.nt[0]; s = 1; }
                            if (s == 1 && cur.n_seg > 2 && tile >= cur.nt[1]) { tile -= cur.nt[1]; s = 2; }
                        }
                        const void* tmap = s == 0 ? cur.tm[0] : s == 1 ? cur.tm[1] : cur.tm[2];
                        const int cs = s == 0 ? cur.cstep[0] : s == 1 ? cur.cstep[1] : cur.cstep[2];
                        const int nb = s == 0 ? cur.bytes[0] : s == 1 ? cur.bytes[1] : cur.bytes[2];
                        const int c0 = ((ce * cs) & ~15) >> 2;   // box start, 16-byte aligned, in 4-byte tensor-map elements
                        // advance the cursor of warp w
                        if (ce + 1 == ept) {
                            c_ce[w] = 0;
                            if (part + 1 == cur.parts) { c_part[w] = 0; c_tile[w]++; } else { c_part[w] = part + 1; }
                        } else {
                            c_ce[w] = ce + 1;
                        }
                        if (!s_wait(rg.empty + 8u * slot, (round & 1u) ^ 1u, s_dead, mp.err, 1000, round * (uint32_t)rg.n_slots + slot)) return;
                        if (sp.no_load) {
                            mbar_arrive(rg.full + 8u * slot);
                        } else {
                            mbar_arrive_expect_tx(rg.full + 8u * slot, (uint32_t)nb);
                            tma_load_2d(rg.base + slot * (uint32_t)kStreamSlotBytes, tmap, c0, tile * kMmaRows, rg.full + 8u * slot);
                        }
                        if (++slot == (uint32_t)rg.n_slots) { slot = 0; round++; }
                    }
                }
            }
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------- consumer side of one GEMV phase
// seq0: number of ring entries this CTA has consumed before this phase (every consumer thread keeps the same count).
template <bool TP, class Pre, class Post>
__device__ __forceinline__ void stream_gemv_cta(const MParams& p, uint8_t* smem, const SRing& rg, uint32_t& seq0, float* s_red,
                                                float (*s_part)[2][2][32], volatile int* s_dead, Pre pre_fn, Post post_fn) {
    pre_fn();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int K = p.K;
    const uint32_t sbase = smem_u32(smem);
    const bool swiglu = p.epi == ME_SWIGLU;
    const int ept = p.s_ept, per_tile = p.s_parts * ept;
    const SDeal d = s_deal(p.s_rot, p.s_ncta, p.s_cbase, p.s_crem, per_tile);
    const int base = d.E / kSW, rem = d.E - base * kSW;
    const int j0 = warp * base + min(warp, rem), n_ent = base + (warp < rem ? 1 : 0), j1 = j0 + n_ent;
    const long long gw = (long long)blockIdx.x * kSW + warp;  // debug stamps only
    MMA_STAMP(0);

    // ---- everything that does not depend on other CTAs happens BEFORE the barrier wait (it used to cost ~0.7 us after it:
    // integer divisions of the cursor, matrix constants, B-operand lane tables) ----
    const XLayout XL = x_layout(K);
    float unscale = 1.0f;
    int pf_s = -1, pf_tile = -1;                          // first tile this warp will finish, and its epilogue operands
    float pf_bias = 0.0f, pf_res = 0.0f, pf_w = 1.0f;
    XSmem sm;   // final after the x staging below (an opaque token is added so that no load of x is hoisted above it)
    sm.p0 = sbase + XL.p0;
    sm.p1 = sbase + XL.p1;
    sm.p2 = sbase + XL.p2;
    sm.sx = sbase + XL.sx;
    sm.x16 = sbase + XL.x16;
    sm.zero = sbase + XL.zero;

    // ---- epilogue of a finished tile: lane L owns row tile*32 + L of segment s (vu: the up row for SwiGLU) ----
    auto epilogue = [&](int s, int tile, float v, float vu) {
        const MSeg& sg = p.seg[s];
        const int j = tile * kMmaRows + lane;
        const bool valid = j < sg.n_rows;
        const bool pf = s == pf_s && tile == pf_tile;   // warp-uniform
        const float e_bias = pf ? pf_bias : (valid && sg.bias) ? sg.bias[j] : 0.0f;
        const float e_res = pf ? pf_res : (valid && p.epi == ME_RESIDUAL) ? p.residual[j] : 0.0f;
        const float e_w = pf ? pf_w : (p.stage_out && p.stage_w && j < p.stage_K) ? p.stage_w[j] : 1.0f;
        v *= unscale;
        float val = v;
        if (swiglu) val = mma_silu(v) * (vu * unscale);
        if (valid) {
            val += e_bias;
            val += e_res;
            if (TP && p.n_peer > 0) {
                for (int r = 0; r < p.n_peer; r++) p.peer_out[r][j] = val;   // partial of a row-parallel GEMV, to every rank (NVLink peer memory)
            } else {
                sg.out[j] = val;
            }
        }
        if (p.stage_out) stage_out32(val, e_w, j, p.stage_K, p.stage_out);
    };
    // logical tile T of the phase -> (segment, tile within the segment)
    auto seg_of = [&](int T, int& s, int& tile) {
        s = 0;
        tile = T;
        if (!swiglu)
            while (s + 1 < p.n_seg && tile >= p.seg[s].n_tiles) { tile -= p.seg[s].n_tiles; s++; }
    };

    // cursor of this warp's run: CTA-local tile, matrix part (SwiGLU: 0 = gate, 1 = up), entry within the row
    int tl = 0, part = 0, ce = 0, s = 0, tile = 0;
    if (n_ent > 0) {
        tl = j0 / per_tile;
        const int r = j0 - tl * per_tile;
        part = r / ept;
        ce = r - part * ept;
        seg_of(d.T0 + tl, s, tile);
    }
    // ring position of the warp's entries: q = seq0 + i * kSW + warp
    uint32_t q = seq0 + (uint32_t)warp;
    uint32_t slot = q % (uint32_t)rg.n_slots, round = q / (uint32_t)rg.n_slots;
    // per-matrix constants of the cursor's position (refreshed when the cursor moves to another matrix)
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
    if (n_ent > 0) load_mat();
    const int sC = p.s_C, n_chunks = p.chunks, n_parts = p.s_parts;
    // the first tile this warp will finish (its start lies in the warp's run): its bias / residual / norm-weight values
    // are fetched right after the barrier, in the shadow of the x staging, instead of in front of the epilogue's stores
    if (n_ent > 0) {
        const int ptl = (j0 + per_tile - 1) / per_tile;
        if (ptl * per_tile < j1) seg_of(d.T0 + ptl, pf_s, pf_tile);
    }

    post_fn();
    MMA_STAMP(1);
    if (pf_s >= 0) {
        const MSeg& sg = p.seg[pf_s];
        const int j = pf_tile * kMmaRows + lane;
        const bool valid = j < sg.n_rows;
        pf_bias = (valid && sg.bias) ? sg.bias[j] : 0.0f;
        pf_res = (valid && p.epi == ME_RESIDUAL) ? p.residual[j] : 0.0f;
        pf_w = (p.stage_out && p.stage_w && j < p.stage_K) ? p.stage_w[j] : 1.0f;
    }

    // ---- x: the staged form left by its producer (flat copy) or f32 -> three int8 planes here ----
    if (p.x_staged) {
        const uint32_t n16 = (XL.zero + 15u) >> 4;
        for (uint32_t i = tid; i < n16; i += kSNT) cp_async16(sbase + 16u * i, p.x_staged + 16u * i);
        cp_async_commit();
        MMA_STAMP(2);
        cp_async_wait<0>();
        if (tid < 64) reinterpret_cast<uint32_t*>(smem + XL.zero)[tid] = 0u;
        cons_sync();
        if (p.norm_w) {
            const float* ssq = reinterpret_cast<const float*>(smem + XL.ssq);
            float tot = 0.0f;
            for (int i = lane; i < (K >> 5); i += 32) tot += ssq[i];
            tot = warp_sum(tot);
            unscale = 1.0f / sqrtf(tot / (float)K + p.eps);
        }
    } else {
        XStage xst;
        // tensor parallel: x = sum of the ranks' partial vectors (+ residual)
        const XSource xsrc{p.x, TP ? p.xsum : nullptr, TP ? p.n_sum : 0, TP ? p.sum_stride : 0, TP ? p.x_res : nullptr, TP ? p.x_full_out : nullptr};
        stage_x_load(xst, xsrc, p.norm_w, K, kSNT);
        MMA_STAMP(2);
        stage_x_finish(xst, xsrc, p.norm_w, K, smem, s_red, kSNT);
        cons_sync();
        unscale = stage_x_unscale(s_red, p.norm_w != nullptr, p.eps, K, kSW);
    }
    MMA_STAMP(3);
    {   // lane tables (lb) hold differences of these addresses, so they stay valid
        const uint32_t tokx = smem_token();
        sm.sx += tokx;
        sm.x16 += tokx;
        sm.zero += tokx;
    }
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f};
    int piece_tl[2] = {-1, -1};   // CTA-local tiles of which this warp holds only a piece (first / last of its run)
    bool first = true;
    for (int i = 0; i < n_ent; i++) {
        const int c0 = ce * sC, nc = min(sC, n_chunks - c0);
        const uint32_t e00 = (uint32_t)c0 * kMmaChunk;
        const uint32_t doff = ((uint32_t)c0 * cbytes) & 15u;   // the box starts 16-byte aligned (Q6_K: any even residue)
        // mbarrier waits are by phase PARITY: before waiting for round r of `full` this warp makes sure the slot's round
        // r - 1 (an entry of ANOTHER warp) has been released, i.e. `full` is in phase r and not still in phase r - 1, where
        // "parity r" would read as already complete.  Unambiguous because no warp is ever more than n_slots entries ahead
        // of the slowest one (n_slots > kSW); round 0 passes at once (fresh barrier, parity 1).
        s_wait(rg.empty + 8u * slot, (round & 1u) ^ 1u, s_dead, p.err, 4000 + warp, q);
        s_wait(rg.full + 8u * slot, round & 1u, s_dead, p.err, 2000 + warp, q);
        if (first) { MMA_STAMP(4); first = false; }
        const uint32_t spb = rg.base + slot * (uint32_t)kStreamSlotBytes + doff + smem_token();
        float ua[4] = {0.f, 0.f, 0.f, 0.f};
        switch (type) {
            case T_Q4_K:
                for (int c = 0; c < nc; c++) unit_k45<false>(spb + (uint32_t)c * cbytes, RS, e00 + (uint32_t)c * kMmaChunk, sm, lb, g, t, ua);
                break;
            case T_Q5_K:
                for (int c = 0; c < nc; c++) unit_k45<true>(spb + (uint32_t)c * cbytes, RS, e00 + (uint32_t)c * kMmaChunk, sm, lb, g, t, ua);
                break;
            case T_Q6_K:
                // block c of the entry starts at doff + 210 c: pick the widest loads its alignment allows (the row pitch is
                // a multiple of 16)
                for (int c = 0; c < nc; c++) {
                    const uint32_t a = spb + (uint32_t)c * cbytes, e0 = e00 + (uint32_t)c * kMmaChunk;
                    const uint32_t al = (doff + (uint32_t)c * cbytes) & 7u;
                    if (al == 0u) unit_q6k<8>(a, RS, e0, 0u, sm, lb, g, t, ua);
                    else if (al == 4u) unit_q6k<4>(a, RS, e0, 0u, sm, lb, g, t, ua);
                    else unit_q6k<2>(a, RS, e0, 0u, sm, lb, g, t, ua);
                }
                break;
            default:
                for (int c = 0; c < nc; c++) {
                    const uint32_t e0 = e00 + (uint32_t)c * kMmaChunk;
                    unit_q80(spb + (uint32_t)c * cbytes, RS, e0, 0u, min(cb, nb_row - (int)(e0 >> 5)), sm, lb, g, t, ua);
                }
                break;
        }
        pin4(ua);   // every shared-memory read of the entry has completed before the slot is handed back
        __syncwarp();
        if (lane == 0) mbar_arrive(rg.empty + 8u * slot);
        q += (uint32_t)kSW;
        slot += (uint32_t)kSW;
        while (slot >= (uint32_t)rg.n_slots) { slot -= (uint32_t)rg.n_slots; round++; }
        if (swiglu && part == 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) au[k] += ua[k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) ag[k] += ua[k];
        }

        // ---- advance the cursor; tile finished (for this warp)? ----
        const bool row_end = ce == ept - 1;
        const bool tile_end = row_end && part == n_parts - 1;
        const bool tile_done = tile_end || (i == n_ent - 1);
        const int cur_tl = tl, cur_s = s, cur_tile = tile;
        ce++;
        if (row_end) {
            ce = 0;
            part++;
            if (tile_end) {
                part = 0;
                tl++;
                if (i + 1 < n_ent) seg_of(d.T0 + tl, s, tile);
            }
            if (i + 1 < n_ent) load_mat();
        }
        if (!tile_done) continue;
        if (i == n_ent - 1) MMA_STAMP(5);
        // reduce the 4 lanes of a row group, then lane L holds logical row L = 8k + n (register k of lanes 4n..4n+3)
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
        const int te0 = cur_tl * per_tile;
        if (j0 <= te0 && te0 + per_tile <= j1) {
            epilogue(cur_s, cur_tile, vg, vu);   // the whole tile is mine
        } else {
            const int ps = (j0 >= te0) ? 0 : 1;   // the tile is my first (0) or starts inside my run (1)
            s_part[warp][ps][0][lane] = vg;
            s_part[warp][ps][1][lane] = vu;
            piece_tl[ps] = cur_tl;
        }
    }
    seq0 += (uint32_t)d.E;

    // ---- merge the pieces of tiles shared between warps, in warp order ----
    cons_sync();
    auto start_of = [&](int w) { return w * base + min(w, rem); };
    auto warp_of = [&](int e) { return e < rem * (base + 1) ? e / (base + 1) : rem + (e - rem * (base + 1)) / max(base, 1); };
#pragma unroll
    for (int ps = 0; ps < 2; ps++) {
        if (piece_tl[ps] < 0) continue;  // warp-uniform
        const int tl = piece_tl[ps], te0 = tl * per_tile;
        const int lo = warp_of(te0), hi = warp_of(te0 + per_tile - 1);
        if (warp != lo) continue;        // the first warp that holds a piece finishes the tile
        float v = 0.f, vu = 0.f;
        for (int w = lo; w <= hi; w++) {
            const int ws = (start_of(w) >= te0) ? 0 : 1;
            v += s_part[w][ws][0][lane];
            vu += s_part[w][ws][1][lane];
        }
        int s, tile;
        seg_of(d.T0 + tl, s, tile);
        epilogue(s, tile, v, vu);
    }
    MMA_STAMP(6);
}

// ---------------------------------------------------------------- grid barrier of the consumer warps
__device__ __forceinline__ void s_grid_arrive(unsigned int* bar, unsigned int& target) {
    cons_sync();
    target += gridDim.x;
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ bool s_grid_wait(unsigned int* bar, unsigned int target, int* err, int* s_flag, volatile int* s_dead) {
    if (threadIdx.x == 0) {
        unsigned int v = 0;
        int ok = 1;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            if (*s_dead || clock64() - t0 > 3000000000LL) {  // ~1.5 s
                ok = 0;
                *s_dead = 1;
                if (atomicExch(err, 2) == 0) { err[1] = 3000; err[2] = (int)blockIdx.x; err[3] = (int)target; }
                break;
            }
        }
        *s_flag = ok;
    }
    cons_sync();
    return *s_flag != 0;
}

// Cross-GPU flag exchange after the local grid barrier (consumer warps only; see tp_exchange in mega.cuh)
__device__ __forceinline__ bool s_tp_exchange(const MegaParams& mp, unsigned int epoch, int* s_flag, volatile int* s_dead) {
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
                if (*s_dead || clock64() - t0 > 3000000000LL) {
                    ok = 0;
                    *s_dead = 1;
                    if (atomicExch(mp.err, 3) == 0) { mp.err[1] = 5000 + r; mp.err[2] = (int)blockIdx.x; mp.err[3] = (int)epoch; }
                    break;
                }
            }
        }
        *s_flag = ok;
    }
    cons_sync();
    return *s_flag != 0;
}

// TP = false is the single-GPU instantiation: no tensor-parallel code in it (it cost 3-4 % of the token when it was a
// run-time branch: registers and instruction footprint)
template <int HD, int GMAX, bool TP>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_decode_kernel(const __grid_constant__ StreamParams sp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bars[2 * kStreamMaxSlots];
    __shared__ float s_red[2 * kSW];
    __shared__ float s_part[kSW][2][2][32];
    __shared__ unsigned int s_ticket;
    __shared__ int s_flag;
    __shared__ int s_dead;
    __shared__ __align__(16) MegaPhase s_phs[3];
    __shared__ float s_rope[HD];
    __shared__ float s_av[kSW];
    __shared__ int s_ai[kSW];

    const MegaParams& mp = sp.mp;
    constexpr int NW = kSW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SRing rg;
    rg.base = smem_u32(smem) + (uint32_t)sp.ring_off;
    rg.full = smem_u32(s_bars);
    rg.empty = rg.full + 8u * (uint32_t)sp.n_slots;
    rg.n_slots = sp.n_slots;
    if (tid == 0) {
        for (int i = 0; i < sp.n_slots; i++) {
            mbar_init(rg.full + 8u * i, 1);
            mbar_init(rg.empty + 8u * i, 1);
        }
        s_dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();   // the only CTA-wide barrier that includes the producer warp

    if (warp == NW) {
        if (lane == 0) stream_producer(sp, rg, &s_dead);
        s_drain(&s_dead);
        return;
    }

    unsigned int target = 0;
    uint32_t seq0 = 0;
    const int n_run = (mp.mode == MEGA_PREFILL) ? mp.n_phases - 1 : mp.n_phases;
    auto fetch_desc = [&](long long gph) {
        if (tid < 32) return;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(mp.phases + (int)(gph % n_run));
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_phs[gph % 3]);
        for (int i = tid - 32; i < (int)(sizeof(MegaPhase) / 4); i += (NW - 1) * 32) dst[i] = src[i];
    };
    fetch_desc(0);
    fetch_desc(1);
    cons_sync();
    long long gph = 0;
    unsigned int tp_n = 0;   // cross-GPU exchanges so far in this launch

    for (int tok = 0; tok < mp.n_tokens; tok++) {
        // ---- embedding (LlamaModel::forward, model/llama.rs:293-306): CTA 0 dequantises row `token`, bit-exactly ----
        if (blockIdx.x == 0) {
            int token = mp.st->token;
            token = min(max(token, 0), mp.vocab - 1);
            const int be = type_block_elems(mp.embd_type), bb = type_block_bytes(mp.embd_type);
            const uint8_t* row = mp.embd + (long long)token * mp.embd_row_bytes;
            for (int i = tid; i < mp.hidden; i += kSNT) {
                const int blk = i / be;
                mp.h[i] = dequant_elem(mp.embd_type, row + (long long)blk * bb, i - blk * be);
            }
            if (tid == 0) {
                const int pn = mp.st->pos_next;
                mp.st->pos_cur = pn;
                mp.st->pos_next = pn + 1;
            }
        }
        bool ok = true;
        bool prev_tp_sync = false;   // the predecessor left partial sums in peer memory
        int stamp = 0;
        auto bar_arrive = [&]() {
            if (TP && prev_tp_sync) {  // the CTA's stores to peer memory must be visible system-wide before the barrier says so
                cons_sync();
                if (tid == 0) asm volatile("fence.acq_rel.sys;" ::: "memory");
            }
            s_grid_arrive(mp.bar, target);
        };
        auto bar_wait = [&]() {
            fetch_desc(gph + 2);
            ok = s_grid_wait(mp.bar, target, mp.err, &s_flag, &s_dead);
            if (TP && prev_tp_sync) ok = s_tp_exchange(mp, mp.tp_epoch0 + (++tp_n), &s_flag, &s_dead) && ok;
            if (mp.dbg && blockIdx.x == 0 && tid == 0) mp.dbg[stamp] = gtimer();
            stamp++;
        };

        for (int ph = 0; ph < n_run; ph++, gph++) {
            const MegaPhase& cur = s_phs[gph % 3];
            if (cur.kind == PH_GEMV) {
                stream_gemv_cta<TP>(cur.gemv, smem, rg, seq0, s_red, s_part, &s_dead, bar_arrive, bar_wait);
            } else {
                attn_stamp(cur.attn, 0);
                bar_arrive();
                bar_wait();
                const AttnParams& ap = cur.attn;
                attn_stamp(ap, 1);
                const int kv_len = *ap.pos + 1;
                if (ph == 1) {   // first attention phase of the token: the rotation angles (Backend::rope, cpu/ops.rs:1216-1337)
                    const float position = (float)(kv_len - 1) / ap.rope_scale;
                    for (int pi = tid; pi < HD / 2; pi += kSNT) {
                        const float theta = position * ap.freq[pi];
                        s_rope[pi] = cosf(theta);
                        s_rope[HD / 2 + pi] = sinf(theta);
                    }
                    cons_sync();
                }
                const int n_items = ap.n_kv * ap.n_splits;
                for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                    const int kh = item / ap.n_splits, split = item - kh * ap.n_splits;
                    attn_decode_item<HD, GMAX, NW>(ap, kh, split, kv_len, reinterpret_cast<float*>(smem), &s_ticket, s_rope);
                    cons_sync();
                }
            }
            if (TP) prev_tp_sync = cur.tp_sync != 0;
        }
        // the barrier that ends the last phase of the token
        gph--;
        bar_arrive();
        bar_wait();
        gph++;
        (void)ok;

        if (mp.mode != MEGA_GREEDY) continue;
        // ---- greedy pick on the device: raw-logit argmax, LAST maximal index wins (src/main.rs:1816-1821) ----
        {
            const int per = (mp.vocab_local + gridDim.x - 1) / gridDim.x;
            const int lo = blockIdx.x * per, hi = min(mp.vocab_local, lo + per);
            float best = -INFINITY;
            int bi = -1;
            for (int i = lo + tid; i < hi; i += kSNT) {
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
            cons_sync();
            if (tid == 0) {
                for (int w = 1; w < NW; w++)
                    if (s_ai[w] >= 0 && (bi < 0 || s_av[w] > best || (s_av[w] == best && s_ai[w] > bi))) { best = s_av[w]; bi = s_ai[w]; }
                mp.cand_val[blockIdx.x] = best;
                mp.cand_idx[blockIdx.x] = bi;
            }
        }
        s_grid_arrive(mp.bar, target);
        s_grid_wait(mp.bar, target, mp.err, &s_flag, &s_dead);
        const unsigned int cand_epoch = mp.tp_epoch0 + ((TP && mp.tp_size > 1) ? ++tp_n : 0u);  // uniform over the grid
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
                if (TP && mp.tp_size > 1) {  // every rank picks the same winner among the per-rank candidates (ties: largest index)
                    bi += mp.tp_rank * mp.vocab_local;
                    if (lane < mp.tp_size) {
                        float* dst = mp.tp_peer_cand[lane] + 2 * mp.tp_rank;
                        asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(best) : "memory");
                        asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + 1), "f"(__int_as_float(bi)) : "memory");
                    }
                    __syncwarp();
                    asm volatile("fence.acq_rel.sys;" ::: "memory");
                    if (lane < mp.tp_size)
                        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mp.tp_peer_flags[lane] + mp.tp_rank), "r"(cand_epoch) : "memory");
                    const long long t0 = clock64();
                    bool okc = true;
                    if (lane < mp.tp_size) {
                        for (;;) {
                            unsigned int v;
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mp.tp_flags + lane) : "memory");
                            if ((int)(v - cand_epoch) >= 0) break;
                            if (clock64() - t0 > 3000000000LL) { okc = false; atomicExch(mp.err, 3); break; }
                        }
                    }
                    __syncwarp();
                    best = -INFINITY;
                    bi = -1;
                    if (okc && lane < mp.tp_size) {
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
            cons_sync();  // CTA 0 embeds the new token at the top of the loop: it must see st->token
        }
    }
    s_drain(&s_dead);
}

}  // namespace b200
