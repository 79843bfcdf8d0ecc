"""Measured logits error of the GEMM prefill / batched decode paths against the oracle (the numbers behind tests/test_gpu_prefill.py's bound)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llama_gguf_b200 as B
import oracle as O
import synth
from synth import rel_err
O.build()
synth.TINY["qwen-kq-tiny"] = dict(arch="qwen2", hidden=512, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=64, ffn=1024,
                                  vocab=600, norm_eps=1e-6, rope_base=1e6, rope_neox=1, bias=True, tied=True)
for preset, mix, n in [("llama-tiny", "Q4_K_M", 48), ("llama-tiny", "Q6_K", 33), ("llama-stream-tiny", "Q4_K_M", 300), ("llama-tiny", "Q4_K_M", 2100),
                       ("tinyllama-tiny", "Q8_0", 64), ("qwen-kq-tiny", "Q5_K_M", 40)]:
    arch, desc, tensors = synth.synth_model(preset, mix, max(384, n + 8))
    gpu = B.GpuOnlyInference(desc, tensors)
    ref = O.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(n, desc["vocab"])
    got = gpu.prefill(prompt)
    want = ref.forward(prompt)
    errs = [rel_err(got, want)]
    tok = O.argmax_last(want)
    for _ in range(4):
        want = ref.forward([tok]); got = gpu.forward(tok); errs.append(rel_err(got, want)); tok = O.argmax_last(want)
    print(f"{preset:20s} {mix:7s} n={n:5d}  prefill logits err {errs[0]:.2e}   decode-after errs {' '.join('%.1e' % e for e in errs[1:])}", flush=True)
    gpu.close()
