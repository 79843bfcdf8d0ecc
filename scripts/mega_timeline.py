"""Per-phase timeline of the per-token megakernel (globaltimer stamps of CTA 0 after each grid barrier).
usage: python scripts/mega_timeline.py [model] [mix] [prompt_len]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama_gguf_b200 as B  # noqa: E402
from llama_gguf_b200.presets import PRESETS, make_desc  # noqa: E402
from llama_gguf_b200.randmodel import random_model  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
mix = sys.argv[2] if len(sys.argv) > 2 else "Q4_K_M"
plen = int(sys.argv[3]) if len(sys.argv) > 3 else 128
preset = PRESETS[model]
desc = make_desc(preset, 8192)
gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, mix, 8192, seed=1236, upload=up))
for i in range(plen):
    gpu.prefill_token((i * 7919 + 1) % desc["vocab"])
L = B.lib()
ctx = gpu._h
buf = (C.c_uint64 * 4096)()
L.b200_debug_mega_timeline(ctx, buf, 4096)  # arm
toks, ms = gpu.decode_greedy(1, 16)
n = L.b200_debug_mega_timeline(ctx, buf, 4096)
t = np.array(buf[:n], dtype=np.float64)
d = np.diff(t) / 1000.0
print(f"{model} {mix}: {ms / 16:.3f} ms/token over 16 tokens; last token: {n - 1} phases, {d.sum():.1f} us between first and last barrier")
names = ["QKV gemv", "rope+attn", "O gemv", "gate/up gemv", "down gemv"]
body = d[:-1].reshape(-1, 5)
for i, nm in enumerate(names):
    print(f"  {nm:14s} mean {body[:, i].mean():7.2f} us  min {body[:, i].min():7.2f}  max {body[:, i].max():7.2f}   x{body.shape[0]} = {body[:, i].sum():8.1f} us")
print(f"  {'vocab head':14s} {d[-1]:7.2f} us")

# ---- inside one GEMV phase: per-warp stamps (start, after pdl point, x loaded + copies issued, x staged,
# first unit landed, last unit computed, exit) ----
stamp_names = ["start", "after wait point", "x loaded+copies issued", "x staged", "first unit landed", "last unit computed", "exit"]
for label, phase in [("O gemv (layer 16)", 16 * 5 + 2), ("gate/up gemv (layer 16)", 16 * 5 + 3), ("down gemv (layer 16)", 16 * 5 + 4), ("QKV gemv (layer 16)", 16 * 5)]:
    if not L.b200_debug_mega_phase(ctx, phase, None, 0):
        continue
    gpu.decode_greedy(1, 2)
    big = (C.c_uint64 * (148 * 16 * 8))()
    m = L.b200_debug_mega_phase(ctx, -1, big, 148 * 16 * 8)
    a = np.array(big[:m], dtype=np.float64).reshape(-1, 8)
    a = a[a[:, 0] > 0]
    t0 = a[:, 0].min()
    print(f"{label}: {a.shape[0]} warps (ns after the first warp's start: min / mean / max)")
    for i, nm in enumerate(stamp_names):
        col = a[:, i][a[:, i] > 0] - t0
        if col.size:
            print(f"   {nm:24s} {col.min():8.0f} / {col.mean():8.0f} / {col.max():8.0f}   ({col.size} warps)")
