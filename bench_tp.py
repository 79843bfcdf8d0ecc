"""Tensor-parallel arm of bench.py: one rank per GPU under torchrun (NCCL for the plumbing: IPC-handle exchange,
barriers, logits gather, max-over-ranks reduction of the timings).  The data path has no NCCL call: partial sums
and argmax candidates travel through NVLink peer memory inside the per-token kernel (csrc/mega.cuh)."""
import json
import os
import time

import numpy as np


def main(args, preset, config, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import bench as BM
    import llama_gguf_b200 as B
    from llama_gguf_b200.parallel import TensorParallelInference
    from llama_gguf_b200.presets import make_desc
    from llama_gguf_b200.randmodel import random_model

    if world != args.gpus:
        BM.log(f"bench.py --gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus} (WORLD_SIZE={world})")
        return 2
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = time.time()
    desc = make_desc(preset, args.ctx)
    repeat = (16 << 20) if args.model in ("llama-3-70b", "mixtral-8x7b") else None   # (host generation time of the 40 GB presets)
    tp = TensorParallelInference(desc, None, device=local_rank,
                                 feeder=lambda up: random_model(preset, args.mix, args.ctx, seed=args.seed, upload=up, repeat_bytes=repeat))
    BM.log(f"rank {rank}: model built and sharded in {time.time() - t0:.1f} s")
    st0 = tp.stats()
    wbytes_local, kvpp_local = st0["weight_bytes_per_token"], st0["kv_bytes_per_pos"]

    for i in range(args.prompt_len):
        tp.prefill_token((i * 7919 + 1) % desc["vocab"])
    tok = BM.host_argmax_last(tp.forward(1))

    # ---- value: device-resident greedy decode, CUDA events on every rank, max over ranks ----
    toks, _ = tp.decode_greedy(tok, args.warmup)
    tok = int(toks[-1])
    sampler = BM.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    l0 = tp.stats()["kernel_launches"]
    kv_len_mid = tp.position() + args.steps // 2
    toks, ms = tp.decode_greedy(tok, args.steps)
    toks_head = toks[:8]
    torch.cuda.synchronize()
    dist.barrier()
    launches = tp.stats()["kernel_launches"] - l0
    ms = max_over_ranks(ms)
    tok = int(toks[-1])
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- e2e: forward() with host token in / full host logits out (slices gathered over NCCL) every step ----
    for _ in range(args.warmup):
        tok = BM.host_argmax_last(tp.forward(tok))
    dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        tok = BM.host_argmax_last(tp.forward(tok))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t1)
    clocks = sampler.stop() if rank == 0 else None

    peaks, peaks_kind = BM.measured_peaks()
    bytes_rank = wbytes_local + kvpp_local * kv_len_mid
    achieved = bytes_rank / (ms_per_step * 1e-3) / 1e9
    if rank == 0:
        roofline = {"bound": "hbm", "kernel": {"stream2": "stream2_decode_kernel", "stream": "stream_decode_kernel", "graph": "gemv_mma_kernel (CUDA graph of per-op kernels)"}.get(tp.path(), "mega_decode_kernel") + " (per rank)", "achieved": achieved, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "peak_kind": f"{peaks_kind} copy bandwidth (MEASURED_PEAKS.json)", "traffic": None,
                    "bytes_per_launch": bytes_rank, "avg_launch_us": ms_per_step * 1e3,
                    "note": ("expert parallel: bytes = the replicated attention / head weights + the selected experts' weights (summed over the GPUs that own them) + KV"
                             if tp.expert_parallel else "one persistent kernel per token and rank; bytes = this rank's weight shard + its KV heads")}
        out = {"metric": BM.METRIC, "value": value, "unit": BM.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config, "clocks": clocks,
               "e2e": {"value": args.steps / e2e_s, "unit": BM.UNIT, "h2d_bytes_per_step": 4 * world,
                       "d2h_bytes_per_step": desc["vocab"] * 4},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
               "weight_bytes_per_token_per_rank": wbytes_local, "kv_bytes_per_token_at_mid_per_rank": kvpp_local * kv_len_mid,
               "greedy_tokens_head": [int(t) for t in toks_head]}
        print(json.dumps(out))
    tp.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0
