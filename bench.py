#!/usr/bin/env python
"""bench.py — decode tok/s of the quantized-forward hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one decoded token (batch 1) through the whole hot path: embedding row,
L x [RMSNorm+QKV dequant-GEMV, RoPE+KV write, GQA attention, O GEMV, RMSNorm+gate/up
GEMV+SwiGLU, down GEMV], final norm + vocab GEMV.  Workload (config.workload): the
configuration the metric is quoted on — Llama-3-8B architecture, Q4_K_M tensor-type mix,
random-init GGUF blocks (synthetic), batch 1, 8K context window, 128-token prompt.

  value  : tokens/s with everything resident in HBM (device-side greedy argmax, CUDA-graph
           replay, no host round trip), timed with CUDA events on the launching stream.
  e2e    : the same metric through the reference-facing call GpuInference::forward —
           token id from host memory in, `vocab` f32 logits to host memory out, host argmax
           (src/main.rs:1811-1822) — wall clock around the K calls.
  roofline     : the per-token kernel (stream_decode_kernel, or mega_decode_kernel for shapes it does not take): weight
           bytes + KV rows one token reads / the CUDA-event time per token of the timed region (the kernel IS the step);
           `gemv_standalone` = the same GEMVs as 129 separate launches, for comparison.
  cpu_baseline : the C++ restatement of the reference's CPU path (oracle/) on this box's cores.

--impl reference times that CPU restatement (the reference is Rust; no rustc in the image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode_tokens_per_sec"
UNIT = "tok/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.lines, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_argmax_last(v):
    """max_by(partial_cmp): the LAST maximal element wins (src/main.rs:1816-1821)."""
    return int(v.size - 1 - np.argmax(v[::-1]))   # (np.argmax returns the FIRST maximum of the reversed view; one pass, no temporaries)


# --------------------------------------------------------------------------- reference arm
def sampled_preset(preset, n_layers_sample):
    p = dict(preset)
    p["n_layers"] = min(preset["n_layers"], n_layers_sample)
    return p


def model_weight_bytes(preset, mix):
    from llama_gguf_b200.presets import tensor_plan
    from llama_gguf_b200.randmodel import BLOCK

    total = 0
    E, k = preset.get("n_experts", 0), preset.get("n_experts_used", 0)
    for name, t, ne in tensor_plan(preset, mix):
        be, bb = BLOCK[t]
        nbytes = int(np.prod(ne)) // be * bb
        if name == "token_embd.weight":
            nbytes = (ne[0] // be * bb) + (nbytes if preset.get("tied") else 0)
        if "_exps." in name:
            nbytes = nbytes // E * k
        total += nbytes
    return total


def run_cpu_reference(preset, mix, ctx_len, prompt_len, steps, warmup, n_layers_sample, seed):
    """Times the oracle's LlamaModel::forward (one decoded token per step) on the box's host cores.

    n_layers_sample = 0: the FULL model, nothing scaled (the --impl reference arm).  n_layers_sample > 0: a layer-bounded sample of
    the workload whose time is scaled by weight bytes to the full model (the bounded cpu_baseline leg of the default run).
    Threads are pinned to the box's core count whatever OMP_NUM_THREADS says (torchrun sets it to 1).  Returns (tok/s, info, ms)."""
    import oracle as O
    from llama_gguf_b200.randmodel import random_model

    O.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.set_num_threads(cores)
    full = n_layers_sample <= 0 or n_layers_sample >= preset["n_layers"]
    ps = dict(preset) if full else sampled_preset(preset, n_layers_sample)
    desc, tensors = random_model(ps, mix, ctx_len, seed=seed)
    ref = O.OracleModel(desc, tensors)
    tok = 1
    n_prompt = min(prompt_len, 4)   # short prompt: the CPU arm measures decode steps, not prefill
    for i in range(n_prompt):
        ref.forward([(i * 7919 + 1) % desc["vocab"]], want_logits=False)
    for _ in range(warmup):
        tok = O.argmax_last(ref.forward([tok]))
    t0 = time.perf_counter()
    for _ in range(steps):
        tok = O.argmax_last(ref.forward([tok]))
    dt = time.perf_counter() - t0
    frac = model_weight_bytes(ps, mix) / model_weight_bytes(preset, mix)
    ms_sample = dt / steps * 1e3
    ms_full = ms_sample / frac
    what = (f"{steps} greedy decode tokens through the full {preset['n_layers']}-layer model (nothing scaled)" if full else
            f"{steps} greedy decode tokens through {ps['n_layers']} of {preset['n_layers']} layers + full vocab head "
            f"({frac * 100:.1f}% of the per-token weight bytes), time scaled by weight bytes to the full model")
    info = {"cores": O.num_threads(), "kind": "port",
            "sample": (what + f", after a {n_prompt}-token prompt; C++ restatement of the reference CPU path (scalar quant dots, threads over "
                       "output rows), hot-path-only embedding (one row per token)"),
            "ms_per_step_sample": ms_sample, "avx512": bool(O.lib().orc_has_avx512())}
    return 1000.0 / ms_full, info, ms_full


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--mix", default="Q4_K_M")
    ap.add_argument("--ctx", type=int, default=8192)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--cpu-layers", type=int, default=4, help="layers in the bounded cpu_baseline sample of the default run")
    ap.add_argument("--ref-layers", type=int, default=0, help="--impl reference: layers timed (0 = the full model, nothing scaled)")
    ap.add_argument("--depth", type=int, default=8000, help="extra line: decode at this KV depth (prompt through b200_prefill; 0 = skip)")
    ap.add_argument("--batch", type=int, default=32, help="extra line: batched decode with this many sequences (0 = skip)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prefill-len", type=int, default=2048, help="tokens of the GEMM-prefill measurement (0 = skip)")
    ap.add_argument("--seed", type=int, default=1236)
    ap.add_argument("--no-speculation", action="store_true", help="e2e: only the plain b200_forward loop (no greedy continuation)")
    ap.add_argument("--kv-format", default=os.environ.get("B200_KV_FORMAT", "f32"), choices=["f32", "int8"],
                    help="KV cache storage (SURVEY 8f row 4): int8 = QuantizedKVCache Int8 rows + one scale per row (per-op decode path)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    from llama_gguf_b200.presets import PRESETS

    preset = PRESETS[args.model]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"{args.model} arch, {args.mix} random-init GGUF blocks, batch-1 greedy decode after a "
                f"{args.prompt_len}-token prompt, {args.ctx}-token context window ({args.kv_format} KV)")
    config = {"workload": workload, "model_arch": args.model, "quant_mix": args.mix, "batch": 1, "context_window": args.ctx,
              "prompt_len": args.prompt_len, "parallelism": (f"ep{args.gpus}" if preset.get("n_experts", 0) else f"tp{args.gpus}") if args.gpus > 1 else "single-gpu",
              "l2_policy": "inputs larger than L2: each step streams the whole weight set (>= 4.6 GB) through the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        v, info, ms_full = run_cpu_reference(preset, args.mix, args.ctx, args.prompt_len, args.steps, args.warmup,
                                             args.ref_layers, args.seed)
        info["value"] = v
        info["unit"] = UNIT
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_full, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": info,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch

    import llama_gguf_b200 as B
    from llama_gguf_b200.randmodel import random_model

    if not torch.cuda.is_available() or B.device_count() == 0:
        log("bench.py: no CUDA device — cuda-b200 has no CPU fallback")
        return 2
    if world > 1 or args.gpus > 1:
        import bench_tp  # tensor-parallel arm lives in its own module

        return bench_tp.main(args, preset, config, rank, world, local_rank)

    torch.cuda.set_device(0)
    t0 = time.time()
    from llama_gguf_b200.presets import make_desc

    desc = make_desc(preset, args.ctx, max(1, args.batch))   # one KV cache per sequence slot (batch extra below)
    repeat = (16 << 20) if args.model in ("llama-3-70b", "mixtral-8x7b") else None   # (host generation time of the 40 GB presets)
    # tensors are streamed straight into the context (host memory stays at one tensor)
    gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, args.mix, args.ctx, seed=args.seed, upload=up, repeat_bytes=repeat),
                             kv_format=args.kv_format)
    log(f"model built and uploaded in {time.time() - t0:.1f} s")
    st0 = gpu.stats()
    wbytes, kvpp = st0["weight_bytes_per_token"], st0["kv_bytes_per_pos"]

    # prompt (untimed)
    for i in range(args.prompt_len):
        gpu.prefill_token((i * 7919 + 1) % desc["vocab"])
    tok = host_argmax_last(gpu.forward(1))

    # ---- value: device-resident greedy decode, CUDA events ----
    toks, _ = gpu.decode_greedy(tok, args.warmup)
    tok = int(toks[-1])
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    torch.cuda.synchronize()
    l0 = gpu.stats()["kernel_launches"]
    kv_len_mid = gpu.position() + args.steps // 2
    toks, ms = gpu.decode_greedy(tok, args.steps)
    toks_head = toks[:8]
    torch.cuda.synchronize()
    launches = gpu.stats()["kernel_launches"] - l0
    tok = int(toks[-1])
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- e2e: GpuInference::forward with host token / host logits ----
    for _ in range(args.warmup):
        tok = host_argmax_last(gpu.forward(tok))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        tok = host_argmax_last(gpu.forward(tok))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t1
    e2e_plain = args.steps / e2e_s
    e2e_val, e2e_mode, spec_stats = e2e_plain, "b200_forward per token, every call waits for its own launch", None
    # the same loop with greedy continuation (b200_ctx_set_speculation): b200_forward picks argmax on the device and launches the
    # next token before it returns; this loop feeds exactly that token back, so the host's turnaround overlaps the next token.
    # Same calls, same host buffers, same bytes per step; the closing synchronize also waits for the one token left in flight.
    if not args.no_speculation:
        try:
            gpu.set_speculation(True)
            if gpu.speculation_stats()["enabled"]:
                for _ in range(args.warmup):
                    tok = host_argmax_last(gpu.forward(tok))
                s0 = gpu.speculation_stats()
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                for _ in range(args.steps):
                    tok = host_argmax_last(gpu.forward(tok))
                torch.cuda.synchronize()
                e2e_spec_s = time.perf_counter() - t1
                s1 = gpu.speculation_stats()
                spec_stats = {"hits": s1["hits"] - s0["hits"], "misses": s1["misses"] - s0["misses"]}
                if spec_stats["misses"] == 0 and args.steps / e2e_spec_s > e2e_plain:
                    e2e_val = args.steps / e2e_spec_s
                    e2e_mode = ("b200_forward per token with greedy continuation (b200_ctx_set_speculation): the device picks argmax and starts the next "
                                "token before the call returns; the caller's token equalled the pick on every step")
            gpu.set_speculation(False)
        except Exception as e:   # the plain number stands
            spec_stats = {"error": str(e)}
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel ----
    peaks, peaks_kind = measured_peaks()
    gms, glaunches, gbytes = gpu.bench_gemv_pass(20)
    # DRAM traffic per token is an ncu measurement (dram__bytes_read + write of one launch / its tokens): it cannot be taken inside a
    # timed run, so it is read from the capture committed for THIS kernel and labelled with its source; null if there is none
    traffic, traffic_src = None, None
    for tname in ("r02_stream2_dram_traffic.json", "gemv_dram_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("kernel", "").startswith(gpu.path() if gpu.path() != "stream2" else "stream2"):
                    traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/" + tname + ": " + tj.get("how", "")
                    break
            except Exception:
                pass
    token_bytes = wbytes + kvpp * kv_len_mid
    mega = launches <= 2  # the per-token megakernel ran: ONE launch covers every decoded token of the timed region
    if mega:
        # the dominant (only) kernel is mega_decode_kernel: algorithmic bytes per token = weights + KV rows read,
        # duration = CUDA-event time per token of the timed region above
        achieved = token_bytes / (ms_per_step * 1e-3) / 1e9
        kname = {"stream2": "stream2_decode_kernel", "stream": "stream_decode_kernel"}.get(gpu.path(), "mega_decode_kernel")
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "peak_kind": f"{peaks_kind} copy bandwidth (MEASURED_PEAKS.json)",
                    "traffic": traffic, "traffic_source": traffic_src, "bytes_per_launch": token_bytes, "avg_launch_us": ms_per_step * 1e3,
                    "launches_per_token": 1, "frac_of_nominal_8TBs": achieved / 8000.0,
                    "note": "one persistent kernel per token (162 phases for Llama-3-8B; stream2 = TMA producer warps + mbarrier ring "
                            "across phase boundaries, stream-K jobs, loader-warp boundary); per launch = per token; the timed region is "
                            f"{args.steps} tokens in ONE launch",
                    "gemv_standalone": {"kernel": "gemv_mma_kernel", "GBps": gbytes / (gms * 1e-3) / 1e9,
                                        "launches_per_token": glaunches, "ms_per_token": gms,
                                        "what": "the same GEMVs as 129 separate PDL-chained launches"}}
    else:
        achieved = gbytes / (gms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "gemv_mma_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "peak_kind": f"{peaks_kind} copy bandwidth (MEASURED_PEAKS.json)",
                    "traffic": traffic, "bytes_per_launch": gbytes / glaunches, "avg_launch_us": gms * 1e3 / glaunches,
                    "launches_per_token": glaunches, "gemv_ms_per_token": gms,
                    "whole_token_frac": token_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- extras: BASELINE configs[2] also names an 8K context and batch 32 ----
    extras = {}
    if args.depth > 0 and args.depth + args.steps + args.warmup + 8 <= args.ctx:
        try:
            gpu.reset()
            gpu.prefill([(i * 7919 + 1) % desc["vocab"] for i in range(args.depth)])   # (GEMM prefill: writes the KV cache)
            td, _ = gpu.decode_greedy(1, args.warmup)
            torch.cuda.synchronize()
            kv_mid_d = gpu.position() + args.steps // 2
            td, ms_d = gpu.decode_greedy(int(td[-1]), args.steps)
            bytes_d = wbytes + kvpp * kv_mid_d
            ach_d = bytes_d / (ms_d / args.steps * 1e-3) / 1e9
            extras["decode_at_depth"] = {"kv_len": int(kv_mid_d), "value": 1000.0 * args.steps / ms_d, "unit": UNIT,
                                         "ms_per_step": ms_d / args.steps, "bytes_per_token": int(bytes_d),
                                         "kv_bytes_per_token": int(kvpp * kv_mid_d), "achieved_GBps": ach_d,
                                         "frac": ach_d / peaks["hbm_gbs"],
                                         "what": f"device-resident greedy decode after a {args.depth}-token prompt (b200_prefill), batch 1"}
        except Exception as e:
            extras["decode_at_depth"] = {"error": str(e)}
    if args.batch > 1:
        try:
            nb = args.batch
            for sq in range(nb):
                gpu.reset(sq)
                gpu.prefill([((i + sq) * 7919 + 1) % desc["vocab"] for i in range(16)], sq)
            seqs = list(range(nb))
            btoks = [1 + sq for sq in range(nb)]
            def bstep(tk):
                lg = gpu.decode_batch(seqs, tk)
                return [host_argmax_last(lg[i]) for i in range(nb)]
            for _ in range(args.warmup):
                btoks = bstep(btoks)
            torch.cuda.synchronize()
            tb = time.perf_counter()
            nsteps_b = max(4, min(args.steps, 32))
            for _ in range(nsteps_b):
                btoks = bstep(btoks)
            torch.cuda.synchronize()
            sb = time.perf_counter() - tb
            ach_b = wbytes / (sb / nsteps_b) / 1e9
            extras["decode_batch"] = {"batch": nb, "value": nb * nsteps_b / sb, "unit": UNIT, "ms_per_step": sb / nsteps_b * 1e3,
                                      "weight_stream_GBps": ach_b, "frac": ach_b / peaks["hbm_gbs"],
                                      "what": f"b200_decode_batch, {nb} sequences, host tokens in / host logits out every step "
                                              "(wall clock): one pass over the weights per step (tcgen05 dequant-GEMM path)"}
            # the same pass with the greedy pick made on the device (what the continuous-batching engine calls): n ids back, not n x vocab logits
            for _ in range(args.warmup):
                btoks = gpu.decode_batch_greedy(seqs, btoks)
            torch.cuda.synchronize()
            tb = time.perf_counter()
            for _ in range(nsteps_b):
                btoks = gpu.decode_batch_greedy(seqs, btoks)
            torch.cuda.synchronize()
            sg = time.perf_counter() - tb
            ach_g = wbytes / (sg / nsteps_b) / 1e9
            extras["decode_batch_greedy"] = {"batch": nb, "value": nb * nsteps_b / sg, "unit": UNIT, "ms_per_step": sg / nsteps_b * 1e3,
                                             "weight_stream_GBps": ach_g, "frac": ach_g / peaks["hbm_gbs"],
                                             "what": f"b200_decode_batch_greedy, {nb} sequences, host tokens in / host token ids out every step "
                                                     "(wall clock): the same GEMM pass, argmax of every row on the device"}
            # continuous batching end to end (b200_batch_*): 2 x nb requests of 16-token prompts, 32 new tokens each, nb slots
            eng = B.BatchedEngine(gpu, max_batch_size=nb, max_seq_len=args.ctx, max_queue_depth=4 * nb, eos_token_id=desc["vocab"] - 1)
            for r in range(2 * nb):
                eng.submit([((i + r) * 7919 + 1) % (desc["vocab"] - 1) for i in range(16)], 32)
            torch.cuda.synchronize()
            tb = time.perf_counter()
            ev = eng.run()
            torch.cuda.synchronize()
            se = time.perf_counter() - tb
            ntok = sum(1 for e in ev if e[0] == "token")
            cnt = eng.counts()
            eng.close()
            extras["batched_engine"] = {"requests": 2 * nb, "slots": nb, "generated_tokens": ntok, "seconds": se, "value": ntok / se, "unit": UNIT,
                                        "steps": cnt["steps"], "decode_rows": cnt["decode_rows"],
                                        "what": "b200_batch_submit / b200_batch_step until every request is done: prompts through b200_prefill, all "
                                                "running sequences of a step in one b200_decode_batch_greedy pass (wall clock, prompts included)"}
        except Exception as e:
            extras.setdefault("decode_batch", {"error": str(e)})
            extras["batch_error"] = str(e)

    # ---- prefill: the whole prompt through the batch entry point (tcgen05 dequant-GEMM, csrc/gemm_umma.cuh) ----
    prefill = None
    if args.prefill_len > 0 and args.prefill_len <= args.ctx:
        try:
            ptoks = [(i * 7919 + 1) % desc["vocab"] for i in range(args.prefill_len)]
            gpu.reset()
            gpu.prefill(ptoks[:64])                      # warm-up (module load, attributes)
            gpu.reset()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            gpu.prefill(ptoks)
            torch.cuda.synchronize()
            pf_s = time.perf_counter() - t2
            from llama_gguf_b200.presets import tensor_plan
            macs = sum(int(np.prod(ne)) for name, tt, ne in tensor_plan(preset, args.mix)
                       if name.startswith("blk.") and name.endswith(".weight") and "norm" not in name)
            tfl = 2.0 * macs * args.prefill_len / pf_s / 1e12
            prefill = {"tokens": args.prefill_len, "seconds": pf_s, "tok_per_s": args.prefill_len / pf_s,
                       "gemm_tflops": tfl, "tensor_peak_tflops": peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")),
                       "tensor_frac": tfl / peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")),
                       "path": "b200_prefill: 2048-token chunks, tcgen05.mma kind::f16 dequant-GEMM + batched RoPE/attention/SwiGLU, "
                               "host tokens in, logits of the last token out (wall clock)",
                       "token_by_token_tok_per_s": 1000.0 / ms_per_step}
        except Exception as e:
            prefill = {"tokens": args.prefill_len, "error": str(e)}

    # ---- CPU baseline (bounded sample, rank 0) ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, info, _ = run_cpu_reference(preset, args.mix, args.ctx, args.prompt_len, args.cpu_steps, 1, args.cpu_layers,
                                           args.seed)
            info["value"], info["unit"] = v, UNIT
            cpu = info
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 4, "d2h_bytes_per_step": desc["vocab"] * 4, "mode": e2e_mode,
                   "plain_value": e2e_plain, "greedy_continuation": spec_stats},
           "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "prefill": prefill, "extras": extras,
           "weight_bytes_per_token": wbytes, "kv_bytes_per_token_at_mid": kvpp * kv_len_mid,
           "greedy_tokens_head": [int(t) for t in toks_head]}
    print(json.dumps(out))
    gpu.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
