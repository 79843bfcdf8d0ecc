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

# ---- inside one GEMV phase (stream2): SM-clock stamps per warp, relative to the moment the CTA's loader saw the grid barrier open ----
if gpu.path() == "stream2" and os.environ.get("PHASES", "1") == "1":
    MHZ = 1965.0
    NC = int(os.environ.get("S2_CONS", "8"))   # consumer warps (kS2Cons); loader = warp NC, producer = warp NC + 1
    lay = int(os.environ.get("LAYER", "16"))
    for label, k in [("QKV gemv", 0), ("O gemv", 2), ("gate/up gemv", 3), ("down gemv", 4)]:
        phase = 1 + lay * 5 + k
        if not L.b200_debug_mega_phase(ctxh, phase, None, 0):
            continue
        gpu.decode_greedy(1, 2)
        big = (C.c_uint64 * (148 * 16 * 8))()
        m = L.b200_debug_mega_phase(ctxh, -1, big, 148 * 16 * 8)
        a = np.array(big[:m], dtype=np.float64).reshape(148, 16, 8)
        ld = a[:, NC, :4]                       # loader: consumers done, arrived, barrier open, x issued
        ref = ld[:, 2][:, None]
        cons = a[:, :NC, :5]
        ok = cons[:, :, 0] > 0
        def st(name, v):
            v = v[np.isfinite(v)]
            if v.size:
                print(f"   {name:46s} min {v.min() / MHZ:7.2f}  mean {v.mean() / MHZ:7.2f}  max {v.max() / MHZ:7.2f} us  ({v.size})")
        print(f"{label} (layer {lay}), per CTA, us relative to 'barrier seen open' by the CTA's loader:")
        st("loader: consumers of previous phase done", (ld[:, 0] - ld[:, 2]))
        st("loader: arrived at grid barrier", (ld[:, 1] - ld[:, 2]))
        st("loader: x copy issued + arrive", (ld[:, 3] - ld[:, 2]))
        rel = np.where(ok[:, :, None], cons - ref[:, :, None], np.nan)
        st("warp: entered phase (previous phase left)", rel[:, :, 0].ravel())
        st("warp: x landed (xfull passed)", rel[:, :, 1].ravel())
        st("warp: first ring entry landed", rel[:, :, 2].ravel())
        st("warp: last entry computed", rel[:, :, 3].ravel())
        st("warp: phase left", rel[:, :, 4].ravel())
        st("CTA: last warp left the phase", np.nanmax(rel[:, :, 4], axis=1))
        st("warp: compute span (first landed -> last computed)", (rel[:, :, 3] - rel[:, :, 2]).ravel())
        st("warp: of which waiting for ring entries", np.where(ok, a[:, :NC, 5], np.nan).ravel())
        st("loader: first tile merged", (a[:, NC, 4] - ld[:, 2]))
        st("loader: last tile's epilogue done", (a[:, NC, 5] - ld[:, 2]))
        pr = a[:, NC + 1, :4]
        st("producer: issue span of the phase's entries", (pr[:, 1] - pr[:, 0]))
        st("producer: of which waiting for free slots", pr[:, 2])
        st("producer: issue span per entry (ns)", (pr[:, 1] - pr[:, 0]) / np.maximum(pr[:, 3], 1) * 1000.0)
        st("producer: non-waiting clocks per entry (x1000)", (pr[:, 1] - pr[:, 0] - pr[:, 2]) / np.maximum(pr[:, 3], 1) * MHZ * 1000.0 / 1000.0)
        L.b200_debug_mega_phase(ctxh, phase, None, 0)  # (re-arming the same phase keeps the buffer; harmless)
