"""Per-phase timeline of the second streamed megakernel (globaltimer stamps of CTA 0's loader warp after each grid barrier).
usage: python scripts/s2_timeline.py [model] [mix] [prompt_len] [ctx]"""
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
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
preset = PRESETS[model]
desc = make_desc(preset, ctx)
gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, mix, ctx, seed=1236, upload=up))
if plen >= 32:
    gpu.forward_batch([(i * 7919 + 1) % desc["vocab"] for i in range(plen)])
else:
    for i in range(plen):
        gpu.prefill_token((i * 7919 + 1) % desc["vocab"])
L = B.lib()
ctxh = gpu._h
buf = (C.c_uint64 * 4096)()
L.b200_debug_mega_timeline(ctxh, buf, 4096)  # arm
toks, ms = gpu.decode_greedy(1, 16)
n = L.b200_debug_mega_timeline(ctxh, buf, 4096)
t = np.array(buf[:n], dtype=np.float64)
print(f"{model} {mix} path {gpu.path()}: {ms / 16:.3f} ms/token over 16 tokens at kv_len ~{plen + 16}")
if gpu.path() == "stream2":
    t = t[1:]   # t[k] = boundary before phase k + 1 (phase 0 = EMBED)
d = np.diff(t) / 1000.0
names = ["QKV gemv", "rope+attn", "O gemv", "gate/up gemv", "down gemv"]
body = d[:-1] if gpu.path() != "stream2" else d[:-1]
body = body[: (len(body) // 5) * 5].reshape(-1, 5)
for i, nm in enumerate(names):
    print(f"  {nm:14s} mean {body[:, i].mean():7.2f} us  min {body[:, i].min():7.2f}  max {body[:, i].max():7.2f}   x{body.shape[0]} = {body[:, i].sum():8.1f} us")
print(f"  {'vocab head':14s} {d[-1]:7.2f} us;  sum of phases {d.sum():.1f} us")
