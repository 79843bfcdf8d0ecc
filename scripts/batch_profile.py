"""One batch-32 decode step (b200_decode_batch) on a 2-layer model with Llama-3-8B shapes: run under ncu for per-kernel times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import llama_gguf_b200 as B
from llama_gguf_b200.presets import PRESETS
from llama_gguf_b200.randmodel import random_model

nb = 32
p = dict(PRESETS["llama-3-8b"]); p["n_layers"] = 2; p["vocab"] = 32768
desc, tensors = random_model(p, "Q4_K_M", 512, max_batch=nb)
gpu = B.GpuOnlyInference(desc, tensors)
rng = np.random.default_rng(1)
for s in range(nb):
    for t in rng.integers(0, desc["vocab"], size=4):
        gpu.prefill_token(int(t), s)
toks = [int(t) for t in rng.integers(0, desc["vocab"], size=nb)]
gpu.decode_batch(list(range(nb)), toks)
gpu.decode_batch(list(range(nb)), toks)
print("ok")
gpu.close()
