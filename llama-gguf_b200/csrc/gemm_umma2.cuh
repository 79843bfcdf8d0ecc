// gemm_umma2.cuh — warp-specialised, persistent tcgen05 / TMEM dequant-GEMM (round 2): same contract as gemm_umma.cuh
//
//   Y[t][j] (+)= sum_k deq(W)[j][k] * X[t][k] (+ bias[j]),   W = N rows of GGUF blocks (Q4_K / Q5_K / Q6_K / Q8_0), X fp16,
//
// (the contraction the reference runs as T GEMVs per matrix: src/model/llama.rs:327-345, src/backend/cuda/gpu_only.rs:776-790;
// block arithmetic: src/tensor/quant/dequant.rs:103-109, 205-356), rebuilt around round 1's measurement that the first kernel
// (128 threads that all dequantise, then all wait for one thread's MMAs) keeps the tensor pipe 10 % busy and streams the
// weights of a 32-row pass at 370 GB/s.  One persistent CTA per SM walks a list of (128 weight rows, TN tokens, K range)
// items; its warps have fixed roles and meet only through mbarriers — there is no CTA-wide barrier in the loop:
//
//   warp 4      PRODUCER   one thread: per 256-element block of K one cp.async.bulk.tensor.2d (SASS UTMALDG) of the RAW
//                          GGUF bytes of 128 weight rows (tensor map over the untouched GGUF layout, box = 128 rows x one
//                          block) into the raw ring, and per 64-element K step one box of TN activation rows x 128 bytes
//                          through a SWIZZLE_128B tensor map over X into the B ring.  It runs ahead across item
//                          boundaries: the rings never drain between tiles.
//   warps 6..   DEQUANT    NG groups of 128 threads (thread = weight row).  Group g takes every NG-th K step: reads its row of
//                          the raw tile from shared memory, dequantises 64 elements (nibbles -> fp16 by a mask, d*sc*q - dmin*m
//                          as ONE f32 FMA on 1024 + q, rounded to fp16) and writes the row's eight 16-byte chunks into a
//                          K-major SWIZZLE_128B stage of the A ring; fence.proxy.async + one mbarrier arrive per warp.
//   warp 5      MMA        one thread: waits for the A and B stage of a step, issues 4 x tcgen05.mma.cta_group::1.kind::f16
//                          (M = 128 rows, N = TN tokens, K = 16; SASS UTCHMMA), tcgen05.commit releases both stages; the
//                          f32 accumulator of an item is one of TWO TMEM buffers, committed to the epilogue at item end.
//   warps 0..3  EPILOGUE   tcgen05.ld (SASS LDTM) of the finished accumulator (warp w = TMEM lanes 32w..), bias / residual
//                          accumulate / split-K partial store, then hands the buffer back: the epilogue of item i overlaps
//                          the main loop of item i + 1.
//
// fp16 operands, f32 accumulation, like gemm_umma.cuh.  Every wait is bounded (~1 s) and reports through p.err.
#pragma once
#include <cuda.h>
#include "gemm_umma.cuh"

namespace b200 {

constexpr int kU2MaxStages = 8;
constexpr int kU2EpiWarps = 4, kU2ProducerWarp = 4, kU2MmaWarp = 5, kU2FirstDeqWarp = 6;

struct Umma2Plan {
    int n_mt, n_nt, n_z;     // row tiles, token tiles, K splits: items = n_mt * n_z * n_nt, token tile fastest
    int sa, sb, sr;          // stages of the A (fp16 weights), B (activations) and raw rings
    int raw_stage;           // bytes per raw stage (128 rows x raw_pitch, rounded up to 128)
    int off_b, off_raw;      // byte offsets of the B and raw rings from the 1024-aligned base (A ring at 0)
    int x3d;                 // activations through the 3-D tensor map (1) or the 2-D fallback {K, T} (0: one load per 64-element slab)
    int dbg;                 // lab only (tools/gemm_lab): 1 = no dequant math, 2 = no raw loads, 4 = no activation loads, 8 = no MMAs,
                             // 64 = no tcgen05.fence in the MMA loop, 128 = plain mbarrier arrives instead of tcgen05.commit (timing only)
    long long* prof;         // lab only: per CTA and role {clocks in the role loop, clocks spent waiting}
};

struct Umma2Bars {
    unsigned long long a_full[kU2MaxStages], a_empty[kU2MaxStages], b_full[kU2MaxStages], b_empty[kU2MaxStages];
    unsigned long long r_full[kU2MaxStages], r_empty[kU2MaxStages], acc_full[2], acc_empty[2];
};

__device__ __forceinline__ void u2_mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void u2_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void u2_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void u2_tma_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void u2_tma_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// bounded wait; `alive` goes false on the first timeout and every later wait of the thread returns at once
__device__ __forceinline__ void u2_wait(uint32_t bar, uint32_t parity, bool& alive, int* err, int code, long long& waited) {
    if (!alive) return;
    {   // fast path: the phase has completed already (no clock reads)
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    const long long t0 = clock64();
    if (!umma_mbar_wait(bar, parity)) {
        alive = false;
        if (err) atomicExch(err, code);
    }
    waited += clock64() - t0;
}

// the nibbles selected by `mask` (bits 0-3 or 4-7 of the two 16-bit halves of w) read as two SUBNORMAL fp16 values n * 2^-24
// (or 16 n * 2^-24): the mask is the whole integer -> float conversion, the power of two goes into the scale
__device__ __forceinline__ float2 u2_nib2(uint32_t w, uint32_t mask) {
    const uint32_t h = w & mask;
    return __half22float2(*reinterpret_cast<const __half2*>(&h));
}

// Q4_K: 64 elements kin .. kin + 63 (kin % 64 == 0) of the row whose raw block starts at `blk` (16-byte aligned, shared memory)
// -> chunks 0..7 of row r of the swizzled fp16 tile.  value = d*sc*q - dmin*m (dequant.rs:205-256) as one f32 FMA
// fma(d*sc * 2^24, q * 2^-24, -dmin*m) (one rounding where the reference has two; invisible after the fp16 rounding).
__device__ __forceinline__ void u2_deq64_q4k(const uint8_t* blk, int kin, uint8_t* sA, int r) {
    const int gp = kin >> 6;
    const uint4 hdr = *reinterpret_cast<const uint4*>(blk);   // d | dmin, scales[12]
    const float d = half_bits_to_float(hdr.x), dmin = half_bits_to_float(hdr.x >> 16);
    // 6-bit scales / mins of sub-blocks 2gp and 2gp + 1 (get_scale_min_k4, dequant.rs:180-199) straight from the header words
    const int b8 = 16 * (gp & 1);
    int s1, m1, s2, m2;
    if (gp < 2) {
        s1 = (hdr.y >> b8) & 63; m1 = (hdr.z >> b8) & 63;
        s2 = (hdr.y >> (b8 + 8)) & 63; m2 = (hdr.z >> (b8 + 8)) & 63;
    } else {
        s1 = ((hdr.w >> b8) & 15) | (((hdr.y >> (b8 + 6)) & 3) << 4);
        m1 = ((hdr.w >> (b8 + 4)) & 15) | (((hdr.z >> (b8 + 6)) & 3) << 4);
        s2 = ((hdr.w >> (b8 + 8)) & 15) | (((hdr.y >> (b8 + 14)) & 3) << 4);
        m2 = ((hdr.w >> (b8 + 12)) & 15) | (((hdr.z >> (b8 + 14)) & 3) << 4);
    }
    const float d1 = __fmul_rn(d, (float)s1) * 16777216.0f, d2 = __fmul_rn(d, (float)s2) * 1048576.0f;   // 2^24; 2^20: high nibbles are 16 q
    const float c1 = -__fmul_rn(dmin, (float)m1), c2 = -__fmul_rn(dmin, (float)m2);
    const uint4* q4 = reinterpret_cast<const uint4*>(blk + 16 + 32 * gp);
    const uint4 qa = q4[0], qb = q4[1];
    const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int c = 0; c < 4; c++) {   // chunk c: elements 8c..8c+7 (low nibbles), chunk 4 + c: elements 32 + 8c.. (high nibbles)
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t w = qw[2 * c + h], w8 = w >> 8;
            const float2 l02 = u2_nib2(w, 0x000F000Fu), l13 = u2_nib2(w8, 0x000F000Fu);    // bytes (0, 2), (1, 3)
            const float2 h02 = u2_nib2(w, 0x00F000F0u), h13 = u2_nib2(w8, 0x00F000F0u);
            lo[2 * h] = umma_pack_h2(fmaf(d1, l02.x, c1), fmaf(d1, l13.x, c1));
            lo[2 * h + 1] = umma_pack_h2(fmaf(d1, l02.y, c1), fmaf(d1, l13.y, c1));
            hi[2 * h] = umma_pack_h2(fmaf(d2, h02.x, c2), fmaf(d2, h13.x, c2));
            hi[2 * h + 1] = umma_pack_h2(fmaf(d2, h02.y, c2), fmaf(d2, h13.y, c2));
        }
        *reinterpret_cast<uint4*>(sA + umma_sw128(r, c)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(sA + umma_sw128(r, 4 + c)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
}

// Q6_K (dequant.rs:321-356), warp-cooperative: blocks are only 2-byte aligned and a raw row pitch of 224 bytes puts the same
// offset of every fourth row on one bank, so thread-per-row reads conflict 8 ways.  Here the 32 lanes walk ONE row together: lane
// (second, i) owns elements e = 32*second + 2i, 2i+1 of the 64-element step (one 16-bit load of ql, one of qh: 64 contiguous
// bytes per warp), and a warp stores one whole 128-byte row of the swizzled tile per instruction.  value = (d * sc) * (q - 32).
__device__ __forceinline__ void u2_deq64_q6k_rows(const uint8_t* raw, int pitch, int kin, uint8_t* sA, int row0, int lane) {
    const int n = kin >> 7, hi = (kin >> 6) & 1, i = lane & 15, second = lane >> 4, qq = 2 * hi + second;
    const int ql_off = 64 * n + 32 * second + 2 * i, qh_off = 128 + 32 * n + 2 * i, sc_off = 192 + 8 * n + (i >> 3) + 2 * qq;
    const int e = 32 * second + 2 * i, sh_lo = 4 * hi, sh_hi = 2 * qq;
    const uint32_t dst = (uint32_t)((e & 7) * 2);
#pragma unroll 4
    for (int rr = 0; rr < 32; rr++) {
        const int r = row0 + rr;
        const uint8_t* blk = raw + r * pitch;
        const uint32_t a = *reinterpret_cast<const unsigned short*>(blk + ql_off), b = *reinterpret_cast<const unsigned short*>(blk + qh_off);
        const float d = half_bits_to_float(*reinterpret_cast<const unsigned short*>(blk + 208));
        const float ds = __fmul_rn(d, (float)(int)*reinterpret_cast<const signed char*>(blk + sc_off));
        const uint32_t q0 = ((a >> sh_lo) & 15u) | (((b >> sh_hi) & 3u) << 4), q1 = ((a >> (8 + sh_lo)) & 15u) | (((b >> (8 + sh_hi)) & 3u) << 4);
        // 0x4B000000 | q is the float 2^23 + q: subtracting 2^23 + 32 gives q - 32 exactly
        const float f0 = __uint_as_float(0x4B000000u | q0) - 8388640.0f, f1 = __uint_as_float(0x4B000000u | q1) - 8388640.0f;
        *reinterpret_cast<uint32_t*>(sA + umma_sw128(r, e >> 3) + dst) = umma_pack_h2(__fmul_rn(ds, f0), __fmul_rn(ds, f1));
    }
}

struct U2Ring {   // ring cursor: stage index + phase parity, advanced without divisions (the role loops are single threads)
    uint32_t i, ph;
    __device__ __forceinline__ void next(uint32_t n) { if (++i == n) { i = 0; ph ^= 1u; } }
};

template <int TN, int NG>
__global__ void __launch_bounds__((kU2FirstDeqWarp + 4 * NG) * 32, 1)
dequant_gemm_umma2_kernel(const __grid_constant__ CUtensorMap xmap, const UmmaParams p, const Umma2Plan pl) {
    extern __shared__ __align__(1024) uint8_t u2_smem[];
    __shared__ Umma2Bars bars;
    __shared__ uint32_t s_tmem;
    __shared__ int s_last;
    constexpr int kCols = 2 * (TN < 32 ? 32 : TN);   // two accumulators
    // activations: one ring stage per 64-element K step (wide token tiles) or per 256-element block = 4 slabs of TN x 128 bytes
    // (token tiles of <= 64 rows: one TMA issue, one barrier round trip and one commit per block instead of four)
    constexpr bool kBBlock = TN <= 64;
    constexpr int kBSlab = TN * 128, kBStage = (kBBlock ? 4 : 1) * kBSlab;
    uint8_t* base = u2_smem + ((1024u - (umma_smem_u32(u2_smem) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + pl.off_b;
    uint8_t* sR = base + pl.off_raw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();   // (programmatic launch: the next kernel's CTAs may be scheduled; they wait for this grid to finish)
    const uint32_t bA_full = umma_smem_u32(bars.a_full), bA_empty = umma_smem_u32(bars.a_empty);
    const uint32_t bB_full = umma_smem_u32(bars.b_full), bB_empty = umma_smem_u32(bars.b_empty);
    const uint32_t bR_full = umma_smem_u32(bars.r_full), bR_empty = umma_smem_u32(bars.r_empty);
    const uint32_t bAcc_full = umma_smem_u32(bars.acc_full), bAcc_empty = umma_smem_u32(bars.acc_empty);
    if (tid == 0) {
        for (int i = 0; i < kU2MaxStages; i++) {
            u2_mbar_init(bA_full + 8u * i, 4);          // one arrive per warp of the dequant group that wrote the stage
            u2_mbar_init(bA_empty + 8u * i, 1);         // tcgen05.commit
            u2_mbar_init(bB_full + 8u * i, 1);          // producer's expect_tx arrive (+ the TMA bytes)
            u2_mbar_init(bB_empty + 8u * i, 1);         // tcgen05.commit
            u2_mbar_init(bR_full + 8u * i, 1);
            u2_mbar_init(bR_empty + 8u * i, 4 * NG);    // every dequant warp has passed the block
        }
        for (int i = 0; i < 2; i++) {
            u2_mbar_init(bAcc_full + 8u * i, 1);        // tcgen05.commit at item end
            u2_mbar_init(bAcc_empty + 8u * i, kU2EpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kU2MmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma_smem_u32(&s_tmem)), "n"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // Programmatic launch: the weights never depend on the predecessor, so only the roles that touch ITS data wait for it
    // (griddepcontrol.wait): the producer before its first activation load -- the raw weight tiles of the first item are already
    // on their way by then -- and the epilogue warps before they read or write Y.  The dequant warps and the MMA thread only see
    // shared memory and TMEM.
    const uint32_t tmem = s_tmem;
    const int n_items = pl.n_mt * pl.n_z * pl.n_nt;
    const uint32_t SA = (uint32_t)pl.sa, SB = (uint32_t)pl.sb, SR = (uint32_t)pl.sr;
    // K range of split z: [z * k_split, min(K, (z + 1) * k_split)), always whole 256-element blocks
    auto item_decode = [&](int item, int& m, int& n, int& kb, int& nblk) {
        n = item % pl.n_nt;
        const int mz = item / pl.n_nt, z = mz % pl.n_z;
        m = mz / pl.n_z;
        kb = p.k_split ? z * p.k_split : 0;
        const int ke = p.k_split ? min(p.K, kb + p.k_split) : p.K;
        nblk = (ke - kb) >> 8;
    };
    bool alive = true;
    long long waited = 0;
    const long long t_begin = pl.prof ? clock64() : 0;
    int role = -1;

    if (warp == kU2ProducerWarp) {
        role = 0;
        if (lane == 0) {
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(p.tmap) : "memory");
            U2Ring rr{0u, 0u}, rb{0u, 0u};
            auto raw_load = [&](int m, int blk) {
                u2_wait(bR_empty + 8u * rr.i, rr.ph ^ 1u, alive, p.err, 11, waited);
                if (pl.dbg & 2) {
                    u2_mbar_arrive(bR_full + 8u * rr.i);
                } else {
                    u2_expect_tx(bR_full + 8u * rr.i, (uint32_t)(p.raw_pitch * kUmmaM));
                    u2_tma_2d(umma_smem_u32(sR) + rr.i * (uint32_t)pl.raw_stage, p.tmap, ((blk * p.raw_bytes) & ~15) >> 2, m * kUmmaM, bR_full + 8u * rr.i);
                }
                rr.next(SR);
            };
            int pre = 0;   // raw blocks of the first item requested before the dependency wait (the raw ring is empty: no wait inside)
            if ((int)blockIdx.x < n_items) {
                int m, n, kb, nblk;
                item_decode(blockIdx.x, m, n, kb, nblk);
                pre = min(nblk, (int)SR);
                for (int b = 0; b < pre; b++) raw_load(m, (kb >> 8) + b);
            }
            pdl_wait();
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int m, n, kb, nblk;
                item_decode(item, m, n, kb, nblk);
                for (int b = 0; b < nblk; b++) {
                    const int blk = (kb >> 8) + b;
                    if (pre > 0) pre--;
                    else raw_load(m, blk);
#pragma unroll 1
                    for (int s = 0; s < (kBBlock ? 1 : 4); s++) {
                        u2_wait(bB_empty + 8u * rb.i, rb.ph ^ 1u, alive, p.err, 12, waited);
                        if (pl.dbg & 4) {
                            u2_mbar_arrive(bB_full + 8u * rb.i);
                        } else {
                            u2_expect_tx(bB_full + 8u * rb.i, (uint32_t)kBStage);
                            const uint32_t dst = umma_smem_u32(sB) + rb.i * (uint32_t)kBStage;
                            if (pl.x3d) {
                                u2_tma_3d(dst, &xmap, 0, n * TN, blk * 4 + s, bB_full + 8u * rb.i);
                            } else {
                                for (int q = 0; q < (kBBlock ? 4 : 1); q++)
                                    u2_tma_2d(dst + (uint32_t)(q * kBSlab), &xmap, blk * 256 + (s + q) * kUmmaK, n * TN, bB_full + 8u * rb.i);
                            }
                        }
                        rb.next(SB);
                    }
                }
            }
        }
    } else if (warp == kU2MmaWarp) {
        role = 1;
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(kUmmaM, TN);
            const uint64_t da0 = umma_desc(umma_smem_u32(sA)), db0 = umma_desc(umma_smem_u32(sB));
            U2Ring ra{0u, 0u}, rb{0u, 0u};
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
                int m, n, kb, nblk;
                item_decode(item, m, n, kb, nblk);
                const uint32_t buf = it & 1u;
                u2_wait(bAcc_empty + 8u * buf, ((it >> 1) & 1u) ^ 1u, alive, p.err, 13, waited);   // the epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem + buf * (uint32_t)TN;
                for (int b = 0; b < nblk; b++) {
                    if (kBBlock) u2_wait(bB_full + 8u * rb.i, rb.ph, alive, p.err, 14, waited);
#pragma unroll 1
                    for (int s = 0; s < 4; s++) {
                        if (!kBBlock) u2_wait(bB_full + 8u * rb.i, rb.ph, alive, p.err, 14, waited);
                        u2_wait(bA_full + 8u * ra.i, ra.ph, alive, p.err, 15, waited);
                        if (!(pl.dbg & 64)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da = da0 + (uint64_t)(ra.i * (uint32_t)(kUmmaM * 128 / 16));
                        const uint64_t db = db0 + (uint64_t)(rb.i * (uint32_t)(kBStage / 16) + (kBBlock ? (uint32_t)s * (uint32_t)(kBSlab / 16) : 0u));
                        if (!(pl.dbg & 8)) {
#pragma unroll
                            for (int kk = 0; kk < kUmmaK / 16; kk++)
                                umma_f16(acc, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc, (b > 0 || s > 0 || kk > 0) ? 1u : 0u);
                        }
                        if (pl.dbg & 128) u2_mbar_arrive(bA_empty + 8u * ra.i); else umma_commit(bA_empty + 8u * ra.i);
                        ra.next(SA);
                        if (!kBBlock) {
                            if (pl.dbg & 128) u2_mbar_arrive(bB_empty + 8u * rb.i); else umma_commit(bB_empty + 8u * rb.i);
                            rb.next(SB);
                        }
                    }
                    if (kBBlock) {
                        if (pl.dbg & 128) u2_mbar_arrive(bB_empty + 8u * rb.i); else umma_commit(bB_empty + 8u * rb.i);
                        rb.next(SB);
                    }
                }
                umma_commit(bAcc_full + 8u * buf);
            }
        }
    } else if (warp >= kU2FirstDeqWarp) {
        const int g = (warp - kU2FirstDeqWarp) >> 2, r = tid - (kU2FirstDeqWarp + 4 * g) * 32;   // group, weight row of the tile
        role = 2 + g;
        // this group's steps are g, g + NG, g + 2 NG, ... of the CTA's global step sequence (every item is whole blocks of 4 steps)
        U2Ring ra{(uint32_t)g, 0u}, rr{0u, 0u};
        bool open = false;                 // a raw block is held (released when the group enters the next one)
        int s = g;                         // step within the current item
        const uint8_t* rraw = sR + r * p.raw_pitch;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int m, n, kb, nblk;
            item_decode(item, m, n, kb, nblk);
            const int steps = nblk * 4;
            int cb = -1;
            for (; s < steps; s += NG) {
                const int b = s >> 2;
                if (b != cb) {
                    if (open) {            // done with the previous block's raw tile
                        __syncwarp();
                        if (lane == 0) u2_mbar_arrive(bR_empty + 8u * rr.i);
                        rr.next(SR);
                    }
                    open = true;
                    cb = b;
                    u2_wait(bR_full + 8u * rr.i, rr.ph, alive, p.err, 16, waited);
                }
                u2_wait(bA_empty + 8u * ra.i, ra.ph ^ 1u, alive, p.err, 17, waited);
                uint8_t* tA = sA + ra.i * (uint32_t)(kUmmaM * 128);
                const int blk = (kb >> 8) + b;   // block of the row
                const uint8_t* rrow = rraw + rr.i * (uint32_t)pl.raw_stage + ((blk * p.raw_bytes) & 15);
                const int kin = (s & 3) * kUmmaK;
                if (pl.dbg & 1) {
                } else if (p.type == T_Q4_K) {
                    u2_deq64_q4k(rrow, kin, tA, r);
                } else if (p.type == T_Q6_K) {
                    u2_deq64_q6k_rows(sR + rr.i * (uint32_t)pl.raw_stage + ((blk * p.raw_bytes) & 15), p.raw_pitch, kin, tA, r & ~31, lane);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(tA + umma_sw128(r, c)) = umma_deq8<true>(p.type, rrow, kin, c);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) u2_mbar_arrive(bA_full + 8u * ra.i);
                ra.i += NG;
                if (ra.i >= SA) { ra.i -= SA; ra.ph ^= 1u; }
            }
            s -= steps;
        }
        if (open) {
            __syncwarp();
            if (lane == 0) u2_mbar_arrive(bR_empty + 8u * rr.i);
        }
    } else {
        role = 5;
        // ---- epilogue warps: warp w owns TMEM lanes 32w..32w+31 = weight rows, 32 token columns per load ----
        pdl_wait();
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
            int m, n, kb, nblk;
            item_decode(item, m, n, kb, nblk);
            const uint32_t buf = it & 1u;
            u2_wait(bAcc_full + 8u * buf, (it >> 1) & 1u, alive, p.err, 18, waited);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int j = m * kUmmaM + tid, tok0 = n * TN;
            const int z = p.k_split ? kb / p.k_split : 0;
            const float bj = (p.bias && j < p.n_rows) ? p.bias[j] : 0.0f;
#pragma unroll 1
            for (int n0 = 0; n0 < TN; n0 += 32) {
                if (tok0 + n0 >= p.T) break;   // warp-uniform
                uint32_t v[32];
                const uint32_t taddr = tmem + buf * (uint32_t)TN + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (j < p.n_rows) {
                    if (p.k_split) {
#pragma unroll
                        for (int q = 0; q < 32; q++)
                            if (tok0 + n0 + q < p.T) p.part[((long long)z * p.T + tok0 + n0 + q) * p.n_rows + j] = __uint_as_float(v[q]);
                    } else if (p.accumulate) {
                        // residual rows first (32 independent loads in flight: a load after a store to the same array would be
                        // serialised by the compiler -- measured 3x on the O projection), then the sums
                        float* yp = p.y + (long long)(tok0 + n0) * p.ldy + j;
                        float old[32];
#pragma unroll
                        for (int q = 0; q < 32; q++) old[q] = (tok0 + n0 + q < p.T) ? __ldcg(yp + (long long)q * p.ldy) : 0.0f;
#pragma unroll
                        for (int q = 0; q < 32; q++)
                            if (tok0 + n0 + q < p.T) yp[(long long)q * p.ldy] = (__uint_as_float(v[q]) + bj) + old[q];
                    } else {
                        float* yp = p.y + (long long)(tok0 + n0) * p.ldy + j;
#pragma unroll
                        for (int q = 0; q < 32; q++)
                            if (tok0 + n0 + q < p.T) yp[(long long)q * p.ldy] = __uint_as_float(v[q]) + bj;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) u2_mbar_arrive(bAcc_empty + 8u * buf);
            if (p.k_split && p.tile_cnt) {
                // split-K without a second kernel: the last CTA to finish this (row tile, token tile) adds the n_z partial tiles in
                // z order (threadFenceReduction pattern: partials -> fence -> counter; the epilogue warps meet on named barrier 1)
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid == 0) {
                    int* cnt = p.tile_cnt + m * pl.n_nt + n;
                    const int last = atomicAdd(cnt, 1) == pl.n_z - 1;
                    if (last) *cnt = 0;        // ready for the next launch
                    s_last = last;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (s_last) {
                    __threadfence();
                    const long long zs = (long long)p.T * p.n_rows;
#pragma unroll 1
                    for (int n0 = 0; n0 < TN && tok0 + n0 < p.T; n0 += 8) {
                        float pv[8][8];        // [z][token]: up to 64 independent loads in flight (at most 8 splits: umma_plan_split)
#pragma unroll
                        for (int zz = 0; zz < 8; zz++)
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                const int tk = tok0 + n0 + q;
                                pv[zz][q] = (zz < pl.n_z && tk < p.T && j < p.n_rows) ? __ldcg(p.part + zz * zs + (long long)tk * p.n_rows + j) : 0.0f;
                            }
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const int tk = tok0 + n0 + q;
                            if (tk < p.T && j < p.n_rows) {
                                float a = 0.0f;
#pragma unroll
                                for (int zz = 0; zz < 8; zz++)
                                    if (zz < pl.n_z) a += pv[zz][q];
                                if (p.bias) a += bj;
                                float* yp = p.y + (long long)tk * p.ldy + j;
                                *yp = p.accumulate ? __ldcg(yp) + a : a;
                            }
                        }
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");   // s_last is rewritten by the next item
            }
        }
    }
    if (pl.prof && lane == 0 && (role < 2 || role == 5 ? true : ((warp - kU2FirstDeqWarp) & 3) == 0) && (role != 5 || warp == 0)) {
        pl.prof[(blockIdx.x * 8 + role) * 2] = clock64() - t_begin;
        pl.prof[(blockIdx.x * 8 + role) * 2 + 1] = waited;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kU2MmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols) : "memory");
    }
}

typedef CUresult (*Umma2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The persistent kernel takes matrices with a raw-tile tensor map (rows and base 16-byte aligned, K % 256 == 0).
inline bool umma2_eligible(const UmmaParams& p) {
    return umma_eligible(p) && p.tmap != nullptr && p.K % 256 == 0 && (!p.k_split || p.k_split % 256 == 0);
}
inline int umma2_tn(int T) { return T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : 256; }

// Ring plan inside `smem_limit` bytes of dynamic shared memory (227 KB opt-in on sm_100a), 1 KB of alignment slack included.
inline bool umma2_plan(const UmmaParams& p, int smem_limit, Umma2Plan& pl, size_t& smem) {
    const int tn = umma2_tn(p.T);
    pl.n_mt = (p.n_rows + kUmmaM - 1) / kUmmaM;
    pl.n_nt = (p.T + tn - 1) / tn;
    pl.n_z = p.k_split ? (p.K + p.k_split - 1) / p.k_split : 1;
    pl.raw_stage = (kUmmaM * p.raw_pitch + 1023) & ~1023;
    const int a = kUmmaM * 128, b = (tn <= 64 ? 4 : 1) * tn * 128;   // B stage: one block (4 slabs) for narrow token tiles, one step otherwise
    int sa = tn >= 128 ? 4 : 6, sb = tn >= 256 ? 4 : tn >= 128 ? 6 : 3, sr = tn >= 128 ? 2 : 3;
    auto total = [&]() { return (size_t)sa * a + (size_t)sb * b + (size_t)sr * pl.raw_stage + 1024; };
    while (total() > (size_t)smem_limit && sa > 3) sa--;
    while (total() > (size_t)smem_limit && sb > 3) sb--;
    while (total() > (size_t)smem_limit && sb > 2 && tn <= 64) sb--;
    while (total() > (size_t)smem_limit && sa > 2) sa--;
    if (total() > (size_t)smem_limit) return false;
    pl.sa = sa; pl.sb = sb; pl.sr = sr;
    pl.off_b = sa * a;
    pl.off_raw = pl.off_b + sb * b;
    smem = total();
    return true;
}

template <int TN, int NG>
inline cudaError_t umma2_launch_tn(const CUtensorMap& xmap, const UmmaParams& p, const Umma2Plan& pl, size_t smem, int n_sm, cudaStream_t st) {
    static bool once[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !once[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dequant_gemm_umma2_kernel<TN, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);   // the static part (barriers) counts against the 227 KB too
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) once[dev] = true;
    }
    const int n_items = pl.n_mt * pl.n_z * pl.n_nt;
    const int grid = std::min(n_items, n_sm);
    if (p.pdl) {   // programmatic stream serialization for both launches (batched decode: ~580 dependent kernels per step)
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3((kU2FirstDeqWarp + 4 * NG) * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, dequant_gemm_umma2_kernel<TN, NG>, xmap, p, pl);
        if (e != cudaSuccess) return e;
        if (pl.n_z > 1 && !p.tile_cnt) {
            const long long n = (long long)p.T * p.n_rows;
            cfg.gridDim = dim3((unsigned)std::min<long long>((n + 255) / 256, 148 * 8)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
            e = cudaLaunchKernelEx(&cfg, umma_reduce_kernel, p, pl.n_z);
            if (e != cudaSuccess) return e;
        }
        return cudaGetLastError();
    }
    dequant_gemm_umma2_kernel<TN, NG><<<grid, (kU2FirstDeqWarp + 4 * NG) * 32, smem, st>>>(xmap, p, pl);
    if (pl.n_z > 1 && !p.tile_cnt) {
        const long long n = (long long)p.T * p.n_rows;
        umma_reduce_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>(p, pl.n_z);
    }
    return cudaGetLastError();
}

// X as a 3-D fp16 tensor {64, T, K / 64} (strides: row pitch ldx, 128 bytes), box = 64 elements (128 bytes, SWIZZLE_128B) x TN rows x
// 1 K step (4 for token tiles of <= 64 rows: one box fills the four slabs of a block); rows past T read as zero.
inline bool umma2_encode_xmap(Umma2EncodeFn encode, CUtensorMap* tm, const UmmaParams& p, int& x3d) {
    const int tn = umma2_tn(p.T);
    {
        const cuuint64_t dims[3] = {(cuuint64_t)kUmmaK, (cuuint64_t)p.T, (cuuint64_t)(p.K / kUmmaK)};
        const cuuint64_t strides[2] = {(cuuint64_t)p.ldx * 2, (cuuint64_t)kUmmaK * 2};
        const cuuint32_t box[3] = {(cuuint32_t)kUmmaK, (cuuint32_t)tn, (cuuint32_t)(tn <= 64 ? 4 : 1)};
        const cuuint32_t estr[3] = {1, 1, 1};
        x3d = 1;
        if (encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)p.x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            return true;
    }
    // fallback (a driver that wants ascending strides): 2-D {K, T}, one box per 64-element slab
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.T};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kUmmaK, (cuuint32_t)tn};
    const cuuint32_t estr[2] = {1, 1};
    x3d = 0;
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)p.x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline void umma2_encode_xmap_2d_only(Umma2EncodeFn encode, CUtensorMap* tm, const UmmaParams& p) {   // lab: force the fallback
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.T};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kUmmaK, (cuuint32_t)umma2_tn(p.T)};
    const cuuint32_t estr[2] = {1, 1};
    encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)p.x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Launch (the caller has set p.tmap / raw_pitch / raw_bytes and, for few-tile passes, p.k_split / p.part).
inline cudaError_t umma2_launch(Umma2EncodeFn encode, const UmmaParams& p, int n_sm, int smem_limit, cudaStream_t st, int ng = 2, int dbg = 0,
                                long long* prof = nullptr) {
    Umma2Plan pl{};
    pl.dbg = dbg;
    pl.prof = prof;
    size_t smem = 0;
    CUtensorMap xmap;
    if (!umma2_plan(p, smem_limit, pl, smem) || !umma2_encode_xmap(encode, &xmap, p, pl.x3d)) return cudaErrorInvalidValue;
    if (dbg & 16) pl.x3d = 0, umma2_encode_xmap_2d_only(encode, &xmap, p);
    const int tn = umma2_tn(p.T);
    if (ng == 4 && tn <= 64 && pl.sa >= 4) {   // narrow token tiles only (the A ring must hold one stage per group)
        if (tn == 32) return umma2_launch_tn<32, 4>(xmap, p, pl, smem, n_sm, st);
        return umma2_launch_tn<64, 4>(xmap, p, pl, smem, n_sm, st);
    }
    if (ng >= 3) {
        if (tn == 32) return umma2_launch_tn<32, 3>(xmap, p, pl, smem, n_sm, st);
        if (tn == 64) return umma2_launch_tn<64, 3>(xmap, p, pl, smem, n_sm, st);
        if (tn == 128) return umma2_launch_tn<128, 3>(xmap, p, pl, smem, n_sm, st);
        return umma2_launch_tn<256, 3>(xmap, p, pl, smem, n_sm, st);
    }
    if (tn == 32) return umma2_launch_tn<32, 2>(xmap, p, pl, smem, n_sm, st);
    if (tn == 64) return umma2_launch_tn<64, 2>(xmap, p, pl, smem, n_sm, st);
    if (tn == 128) return umma2_launch_tn<128, 2>(xmap, p, pl, smem, n_sm, st);
    return umma2_launch_tn<256, 2>(xmap, p, pl, smem, n_sm, st);
}

}  // namespace b200
