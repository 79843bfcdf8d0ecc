"""Stamps inside one attention phase of the megakernel (thread 0 of every CTA): phase entered, barrier open, RoPE done,
KV loop done, partial written, ticket taken, merge done.  usage: python scripts/attn_timeline.py [layer]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama_gguf_b200 as B  # noqa: E402
from llama_gguf_b200.presets import PRESETS, make_desc  # noqa: E402
from llama_gguf_b200.randmodel import random_model  # noqa: E402

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 16
preset = PRESETS["llama-3-8b"]
desc = make_desc(preset, 8192)
gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, "Q4_K_M", 8192, seed=1236, upload=up))
for i in range(128):
    gpu.prefill_token((i * 7919 + 1) % desc["vocab"])
L = B.lib()
ctx = gpu._h
assert L.b200_debug_mega_phase(ctx, layer * 5 + 1, None, 0)
gpu.decode_greedy(1, 2)
big = (C.c_uint64 * (148 * 8 * 8))()
m = L.b200_debug_mega_phase(ctx, -1, big, 148 * 8 * 8)
a = np.array(big[:148 * 8], dtype=np.float64).reshape(148, 8)
t0 = a[:, 0][a[:, 0] > 0].min()
names = ["phase entered", "barrier open", "rope done", "kv loop done", "partial written", "ticket taken", "merge done"]
for i, nm in enumerate(names):
    col = a[:, i][a[:, i] > 0] - t0
    if col.size:
        print(f"   {nm:18s} {col.min():8.0f} / {col.mean():8.0f} / {col.max():8.0f}   ({col.size} CTAs)")
