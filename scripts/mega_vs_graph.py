"""Debug aid: per-token logits of the megakernel path against the CUDA-graph path on a tiny model."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import llama_gguf_b200 as B  # noqa: E402
import synth  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else "tinyllama-tiny"
mix = sys.argv[2] if len(sys.argv) > 2 else "Q8_0"
arch, desc, tensors = synth.synth_model(preset, mix, 96)
os.environ["B200_MEGA"] = "0"
g = B.GpuOnlyInference(desc, tensors)
os.environ["B200_MEGA"] = "1"
m = B.GpuOnlyInference(desc, tensors)
toks = synth.prompt_tokens(20, desc["vocab"])
for i, t in enumerate(toks):
    if i % 2 == 0:
        a, b = g.forward(t), m.forward(t)
        print(i, "forward  rel err", synth.rel_err(b, a), "nan" if np.isnan(b).any() else "")
    else:
        g.prefill_token(t)
        m.prefill_token(t)
        print(i, "prefill")

print("--- consecutive prefills, then forward")
for n in (1, 2, 3, 5, 8, 15):
    g.reset()
    m.reset()
    for t in toks[:n]:
        g.prefill_token(t)
        m.prefill_token(t)
    a, b = g.forward(toks[n]), m.forward(toks[n])
    print(n, "prefills then forward: rel err", synth.rel_err(b, a), "nan" if np.isnan(b).any() else "")
