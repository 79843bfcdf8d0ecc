"""Batched decode (BASELINE configs[2]: Llama-3-8B Q4_K_M, batch 32): tokens/s of b200_decode_batch through the tcgen05
dequant-GEMM pass against the same 32 sequences decoded one after the other.  usage: python scripts/batch_bench.py [B] [steps]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import llama_gguf_b200 as B
from llama_gguf_b200.presets import PRESETS, make_desc
from llama_gguf_b200.randmodel import random_model

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
preset = PRESETS["llama-3-8b"]
desc = make_desc(preset, 2048, nb)
gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, "Q4_K_M", 2048, seed=1236, max_batch=nb, upload=up))
slots = list(range(nb))
rng = np.random.default_rng(1)
for s in slots:
    gpu.prefill([int(t) for t in rng.integers(0, desc["vocab"], size=64)], s) if hasattr(gpu, "prefill_seq") else [gpu.prefill_token(int(t), s) for t in rng.integers(0, desc["vocab"], size=8)]
res = {}
for mode, env in (("gemm", "8"), ("sequential", "100000")):
    gpu.batch_gemm_min = None
    toks = [int(t) for t in rng.integers(0, desc["vocab"], size=nb)]
    if mode == "sequential":
        t0 = time.perf_counter()
        for _ in range(max(2, steps // 4)):
            for s in slots:
                gpu.forward(toks[s], s)
        dt = (time.perf_counter() - t0) / max(2, steps // 4)
    else:
        gpu.decode_batch(slots, toks)
        t0 = time.perf_counter()
        for _ in range(steps):
            gpu.decode_batch(slots, toks)
        dt = (time.perf_counter() - t0) / steps
    res[mode] = {"ms_per_step": dt * 1e3, "tok_per_s": nb / dt}
print(json.dumps({"batch": nb, "model": "llama-3-8b Q4_K_M", **res}))
gpu.close()
