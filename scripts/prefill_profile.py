"""One GEMM prefill of N tokens on a 2-layer model with Llama-3-8B shapes (random GGUF blocks): run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel times.  usage: python scripts/prefill_profile.py [N] [mix]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama_gguf_b200 as B
from llama_gguf_b200.presets import PRESETS, make_desc
from llama_gguf_b200.randmodel import random_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mix = sys.argv[2] if len(sys.argv) > 2 else "Q4_K_M"
p = dict(PRESETS["llama-3-8b"]); p["n_layers"] = 2; p["vocab"] = 4096
desc, tensors = random_model(p, mix, 4096)
gpu = B.GpuOnlyInference(desc, tensors)
toks = [(i * 7919 + 1) % desc["vocab"] for i in range(n)]
gpu.prefill(toks[:64]); gpu.reset()
import time; t = time.perf_counter(); gpu.prefill(toks); print("prefill", n, "tokens:", time.perf_counter() - t, "s")
gpu.close()
