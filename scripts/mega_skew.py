"""Which CTAs are late?  Per-warp stamps of several GEMV phases of the megakernel: per-CTA time of 'last unit computed' and
'exit' relative to the phase mean, and the correlation of that lateness between phases (systematic vs random skew).
usage: python scripts/mega_skew.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama_gguf_b200 as B  # noqa: E402
from llama_gguf_b200.presets import PRESETS, make_desc  # noqa: E402
from llama_gguf_b200.randmodel import random_model  # noqa: E402

preset = PRESETS["llama-3-8b"]
desc = make_desc(preset, 8192)
gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, "Q4_K_M", 8192, seed=1236, upload=up))
for i in range(128):
    gpu.prefill_token((i * 7919 + 1) % desc["vocab"])
L = B.lib()
ctx = gpu._h
NW = 8
res = {}
for label, phase in [("gate/up L16", 16 * 5 + 3), ("gate/up L17", 17 * 5 + 3), ("down L16", 16 * 5 + 4), ("QKV L16", 16 * 5), ("O L16", 16 * 5 + 2)]:
    if not L.b200_debug_mega_phase(ctx, phase, None, 0):
        print("no debug phase"); break
    gpu.decode_greedy(1, 2)
    big = (C.c_uint64 * (148 * 16 * 8))()
    m = L.b200_debug_mega_phase(ctx, -1, big, 148 * 16 * 8)
    a = np.array(big[:m], dtype=np.float64).reshape(-1, 8)[:148 * NW].reshape(148, NW, 8)
    t0 = a[:, :, 0][a[:, :, 0] > 0].min()
    a = a - t0
    last = a[:, :, 5].max(axis=1)   # per CTA: last unit computed
    ex = a[:, :, 6].max(axis=1)     # per CTA: exit
    st = a[:, :, 0].min(axis=1)
    aw = a[:, :, 1].max(axis=1)
    res[label] = (last, ex)
    print(f"{label}: CTA start {st.min():.0f}/{st.mean():.0f}/{st.max():.0f}  barrier open {aw.min():.0f}/{aw.mean():.0f}/{aw.max():.0f}  "
          f"last unit {last.min():.0f}/{last.mean():.0f}/{last.max():.0f}  exit {ex.min():.0f}/{ex.mean():.0f}/{ex.max():.0f} ns")
    order = np.argsort(ex)
    print("   latest CTAs (exit):", [(int(c), int(ex[c] - ex.mean())) for c in order[-8:]])
    print("   earliest CTAs     :", [(int(c), int(ex[c] - ex.mean())) for c in order[:6]])
keys = list(res)
for i in range(len(keys)):
    for j in range(i + 1, len(keys)):
        c1 = np.corrcoef(res[keys[i]][0], res[keys[j]][0])[0, 1]
        c2 = np.corrcoef(res[keys[i]][1], res[keys[j]][1])[0, 1]
        print(f"corr {keys[i]} ~ {keys[j]}: last-unit {c1:.2f} exit {c2:.2f}")
