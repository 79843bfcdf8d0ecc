// misc.cuh — small memory-bound kernels around the GEMVs: embedding row, argmax,
// MoE routing, and the element-wise / normalisation ops of the Backend surface.
#pragma once
#include "common.cuh"
#include "quant.cuh"

namespace b200 {

// Token state of one sequence slot on the device.  pos_cur is what the layer kernels
// read; the embedding kernel (first kernel of every token) advances it, so one CUDA
// graph can be replayed for every token of a sequence.
struct SeqState {
    int pos_cur;     // position of the token being processed
    int pos_next;    // position the next token will get
    int token;       // id of the token to process (host write or device argmax)
    int n_generated; // tokens appended to `generated` by the device argmax
};

// embedding lookup: row `token` of token_embd, dequantised bit-exactly
// (LlamaModel::forward, src/model/llama.rs:293-306; cuda/gpu_only.rs:849-858 does it on the host).
__global__ void embed_kernel(int type, const uint8_t* __restrict__ table, long long row_bytes, int hidden,
                             SeqState* st, float* __restrict__ x, int vocab) {
    pdl_launch_dependents();
    pdl_wait();
    int tok = st->token;
    tok = min(max(tok, 0), vocab - 1);
    const int be = type_block_elems(type), bb = type_block_bytes(type);
    const uint8_t* row = table + (long long)tok * row_bytes;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hidden; i += gridDim.x * blockDim.x) {
        int blk = i / be;
        x[i] = dequant_elem(type, row + (long long)blk * bb, i - blk * be);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->pos_cur = st->pos_next;
        st->pos_next = st->pos_next + 1;
    }
}

// Greedy pick on the device: raw-logit argmax, LAST maximal index wins
// (bench loop src/main.rs:1816-1821: max_by keeps the last of equal maxima).
// Single CTA; writes the next token into the slot state and the output ring.
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, int n, SeqState* st,
                                                      int* __restrict__ generated, int max_generated) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    float best = -INFINITY;
    int bi = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = logits[i];
        if (v >= best || bi < 0) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        best = threadIdx.x < nw ? s_v[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < nw ? s_i[threadIdx.x] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) {
            st->token = bi;
            int g = st->n_generated;
            if (g < max_generated) generated[g] = bi;
            st->n_generated = g + 1;
        }
    }
}

// The same pick for every row of a [rows][n] logits matrix (batched decode): one CTA per row, out[row] = LAST maximal index.
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int n, int* __restrict__ out) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    const float* row = logits + (size_t)blockIdx.x * n;
    float best = -INFINITY;
    int bi = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = row[i];
        if (v >= best || bi < 0) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        best = threadIdx.x < nw ? s_v[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < nw ? s_i[threadIdx.x] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) out[blockIdx.x] = bi;
    }
}

// MoeRouter::route (src/model/moe.rs:128-198) with normalize=false: logits = W_r h (f32),
// stable descending top-k (ties -> lowest expert index), softmax over the k selected.
// One CTA; warp e computes logit e.  h is the (already RMS-normalised) FFN input.
struct RouteParams {
    const float* x;        // [hidden] un-normalised hidden state
    const float* norm_w;   // ffn_norm weight (fused RMSNorm, same as the GEMV prologue)
    float eps;
    const float* w_router; // [n_experts][hidden] f32
    int hidden, n_experts, top_k;
    int* sel;              // out [top_k]
    float* wt;             // out [top_k]
    unsigned int* epoch;   // optional (expert parallel): bumped once per call -- the layer's exchange epoch, the same on every GPU
};
__global__ void __launch_bounds__(256) moe_route_kernel(const RouteParams p) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[8];
    __shared__ float logits[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < p.hidden; i += blockDim.x) { float v = p.x[i]; ss += v * v; }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < 8; w++) tot += red[w];
    const float inv = 1.0f / sqrtf(tot / (float)p.hidden + p.eps);
    for (int e = warp; e < p.n_experts; e += 8) {
        const float* wr = p.w_router + (size_t)e * p.hidden;
        float d = 0.0f;
        for (int i = lane; i < p.hidden; i += 32) d = fmaf((p.x[i] * inv) * p.norm_w[i], wr[i], d);
        d = warp_sum(d);
        if (lane == 0) logits[e] = d;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.epoch) *p.epoch += 1u;
        unsigned long long taken = 0ull;
        float sel_l[8];
        float mx = -INFINITY;
        for (int k = 0; k < p.top_k; k++) {
            int best = -1;
            for (int e = 0; e < p.n_experts; e++) {
                if ((taken >> e) & 1ull) continue;
                if (best < 0 || logits[e] > logits[best]) best = e;  // strict >: lowest index wins ties
            }
            taken |= 1ull << best;
            p.sel[k] = best;
            sel_l[k] = logits[best];
            mx = fmaxf(mx, sel_l[k]);
        }
        float sum = 0.0f;
        for (int k = 0; k < p.top_k; k++) sum += expf(sel_l[k] - mx);
        for (int k = 0; k < p.top_k; k++) p.wt[k] = expf(sel_l[k] - mx) / sum;
    }
}

// ---- Backend element-wise surface (kernels.rs:48-95 equivalents; cpu/ops.rs) ----
enum : int { EW_ADD = 0, EW_MUL = 1, EW_SCALE = 2, EW_SILU = 3, EW_GELU = 4 };
__global__ void elementwise_kernel(int op, const float* a, const float* b, float s, float* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x = a[i], r;
        switch (op) {
            case EW_ADD: r = __fadd_rn(x, b[i]); break;
            case EW_MUL: r = __fmul_rn(x, b[i]); break;
            case EW_SCALE: r = __fmul_rn(x, s); break;
            case EW_SILU: r = x / (1.0f + expf(-x)); break;
            default: {  // cpu/ops.rs:326-346
                float inner = __fmul_rn(0.7978846f, __fadd_rn(x, __fmul_rn(__fmul_rn(__fmul_rn(0.044715f, x), x), x)));
                r = __fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, tanhf(inner)));
            }
        }
        out[i] = r;
    }
}

// Backend::softmax over one row (cpu/ops.rs:348-384).  One CTA per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* x, float* out, int n) {
    __shared__ float red[8];
    const float* xr = x + (size_t)blockIdx.x * n;
    float* orow = out + (size_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, xr[i]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float e = expf(xr[i] - mx);
        orow[i] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < 8; w++) tot += red[w];
    const float inv = 1.0f / tot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) orow[i] *= inv;
}

// Backend::rms_norm (cpu/ops.rs:392-422 -> simd.rs:847-899).  One CTA per row.
__global__ void __launch_bounds__(256) rms_norm_rows_kernel(const float* x, const float* w, float eps, float* out, int n) {
    __shared__ float red[8];
    const float* xr = x + (size_t)blockIdx.x * n;
    float* orow = out + (size_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ss = fmaf(xr[i], xr[i], ss);
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.0f;
    for (int k = 0; k < 8; k++) tot += red[k];
    const float inv = 1.0f / sqrtf(tot / (float)n + eps);
    for (int i = threadIdx.x; i < n; i += blockDim.x) orow[i] = __fmul_rn(__fmul_rn(xr[i], inv), w[i]);
}

// Backend::vec_mat for arbitrary k (cpu/ops.rs:959-1005): warp per output row.
__global__ void vec_mat_f32_kernel(const float* a, const float* w, float* out, int k, int n) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const float* wr = w + (size_t)row * k;
    float s = 0.0f;
    for (int i = lane; i < k; i += 32) s = fmaf(a[i], wr[i], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// Backend::rope on host-provided q/k (per-op surface): same math as rope_kv_kernel.
__global__ void rope_inplace_kernel(float* q, float* k, const float* freq, int n_heads, int n_kv, int hd, int pos,
                                    float rope_scale, int neox) {
    const int half = hd >> 1;
    const int n_pairs = (n_heads + n_kv) * half;
    const float position = (float)pos / rope_scale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        const int head = i / half, pi = i - head * half;
        const float theta = position * freq[pi];
        const float c = cosf(theta), s = sinf(theta);
        float* d = head < n_heads ? q + (size_t)head * hd : k + (size_t)(head - n_heads) * hd;
        const int i0 = neox ? pi : 2 * pi;
        const int i1 = neox ? pi + half : 2 * pi + 1;
        const float x0 = d[i0], x1 = d[i1];
        d[i0] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        d[i1] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
    }
}

// Backend::matmul (src/backend/cpu/ops.rs:429-487): out[m][n] = a[m][k] @ b[k][n], all row-major f32; one thread per output element,
// k summed in order with separate multiply and add like matmul_simple (no FMA contraction in the reference).  Compatibility surface.
__global__ void matmul_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int m, int k, int n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * n) return;
    const int i = (int)(idx / n), j = (int)(idx - (long long)i * n);
    float sum = 0.0f;
    for (int kk = 0; kk < k; kk++) sum = __fadd_rn(sum, __fmul_rn(a[(size_t)i * k + kk], b[(size_t)kk * n + j]));
    out[idx] = sum;
}


// ---------------------------------------------------------------- expert parallel (Mixtral across the GPUs of one box)
// Replaces the reference's rayon loop over the selected experts on ONE host (src/model/moe.rs:352-361) and its per-token H2D copies
// (src/backend/cuda/gpu_only.rs:1834-1857).  Every GPU runs attention and the router (replicated: the same selection everywhere);
// expert e lives on GPU e / (E / P).  After the expert GEMVs of a layer:
//   ep_push_kernel    : for every slot whose expert is local, the weighted output y_s = w_s * down_s(...) goes to EVERY GPU's buffer as
//                       8-byte (value, epoch) packets over NVLink peer memory -- no fence, no flag: a packet is valid when its epoch is
//                       this layer's; thread 0 also leaves this GPU's arrival mark with every peer (lock step: nobody runs two layers
//                       ahead of a GPU that owned nothing, so the two packet buffers can alternate);
//   ep_combine_kernel : polls the k slots' packets (and the arrival marks) in local memory and writes xa = ((0 + y_0) + y_1 ...) + h,
//                       the order of moe.rs:363-368 + layers.rs:1235-1241 -- bit-identical to the single-GPU path.
// At batch 1 the dispatch is free (the router runs everywhere); the combine is the only exchange.
struct EpParams {
    const int* sel;               // [top_k] selected experts (device, written by moe_route_kernel)
    const unsigned int* epoch;    // device counter: moe_route_kernel bumps it once per layer (the same count on every GPU)
    int top_k, hidden, rank, world, experts_per_rank;
    const float* y;               // [top_k][hidden] local weighted expert outputs (only the local slots are meaningful)
    uint2* peer_ll[8];            // every GPU's [2][8][hidden] packet buffers as mapped here
    unsigned int* peer_mark[8];   // every GPU's [8] arrival marks as mapped here
    const uint2* ll;              // this GPU's packet buffers
    const unsigned int* mark;     // this GPU's arrival marks
    const float* h;               // residual
    float* out;                   // [hidden]
    int* err;
};
__global__ void __launch_bounds__(256) ep_push_kernel(const EpParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned int ep = *p.epoch;
    const size_t buf = (size_t)(ep & 1u) * 8 * p.hidden;
    for (int s = 0; s < p.top_k; s++) {
        if (p.sel[s] / p.experts_per_rank != p.rank) continue;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.hidden; j += gridDim.x * blockDim.x) {
            const unsigned int v = __float_as_uint(p.y[(size_t)s * p.hidden + j]);
            for (int r = 0; r < p.world; r++)
                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p.peer_ll[r] + buf + (size_t)s * p.hidden + j), "r"(v), "r"(ep) : "memory");
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int r = 0; r < p.world; r++) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p.peer_mark[r] + p.rank), "r"(ep) : "memory");
}
__global__ void __launch_bounds__(256) ep_combine_kernel(const EpParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned int ep = *p.epoch;
    const size_t buf = (size_t)(ep & 1u) * 8 * p.hidden;
    const long long t0 = clock64();
    bool dead = false;
    if (threadIdx.x < p.world) {   // lock step: every GPU has reached this layer's push
        unsigned int v;
        for (;;) {
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p.mark + threadIdx.x) : "memory");
            if ((int)(v - ep) >= 0) break;
            if (clock64() - t0 > 4000000000LL) { dead = true; break; }
        }
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.hidden; j += gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < p.top_k; s++) {
            unsigned int v, e;
            for (;;) {
                asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(p.ll + buf + (size_t)s * p.hidden + j) : "memory");
                if (e == ep) break;
                if (clock64() - t0 > 4000000000LL) { dead = true; break; }
            }
            acc += __uint_as_float(v);
        }
        p.out[j] = acc + p.h[j];
    }
    if (dead && p.err) atomicExch(p.err, 8);
    __syncthreads();   // (the arrival-mark wait of the first threads holds the whole CTA)
}

}  // namespace b200
