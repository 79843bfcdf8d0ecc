"""Synthetic random-init models for the parity tests (SURVEY.md §8d recipe).

Weights are drawn N(0, sigma^2) and quantised PER TENSOR TYPE with the restated reference
quantisers (oracle.quantize -> src/tensor/quant/dequant.rs:374-999), norm weights are
1 + N(0, 0.1^2), qwen2 biases N(0, 0.02^2), MoE routers F32.  Test infrastructure: imports
the oracle, so nothing under llama-gguf_b200/ may import this file.
"""
import numpy as np

import oracle as O
from llama_gguf_b200.presets import F32, PRESETS, make_desc, tensor_plan

# Reduced shapes that keep every structural feature of the named architectures but let the
# CPU oracle finish in seconds.
TINY = {
    # llama arch, hd=128, G=4, Q4_K_M-style mixes, K-quant everywhere
    "llama-tiny": dict(arch="llama", hidden=1024, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=128, ffn=2048,
                       vocab=768, norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False),
    # qwen2 arch: NeoX RoPE, biases, tied embeddings, hidden % 256 != 0 (Q5_0/Q8_0 fallbacks), G=7
    "qwen-tiny": dict(arch="qwen2", hidden=448, n_layers=3, n_heads=7, n_kv_heads=1, head_dim=64, ffn=768,
                      vocab=1000, norm_eps=1e-6, rope_base=1e6, rope_neox=1, bias=True, tied=True),
    # tinyllama-like: hd=64, G=8
    "tinyllama-tiny": dict(arch="llama", hidden=512, n_layers=2, n_heads=8, n_kv_heads=1, head_dim=64, ffn=1536,
                           vocab=512, norm_eps=1e-5, rope_base=1e4, rope_neox=0, bias=False, tied=False),
    # shapes the streamed megakernel takes (csrc/stream.cuh): every weight row a 16-byte multiple, i.e. K % 2048 == 0 for
    # the Q6_K tensors of a *_K_M mix; vocab not a multiple of 32 (ragged last tile, zero-filled by the TMA unit)
    "llama-stream-tiny": dict(arch="llama", hidden=2048, n_layers=2, n_heads=16, n_kv_heads=4, head_dim=128, ffn=4096,
                              vocab=1000, norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False),
    "tinyllama-stream-tiny": dict(arch="llama", hidden=2048, n_layers=2, n_heads=32, n_kv_heads=4, head_dim=64, ffn=2048,
                                  vocab=520, norm_eps=1e-5, rope_base=1e4, rope_neox=0, bias=False, tied=False),
    # mixtral-like: 4 experts top-2
    "mixtral-tiny": dict(arch="llama", hidden=512, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=64, ffn=512,
                         vocab=512, norm_eps=1e-5, rope_base=1e6, rope_neox=0, bias=False, tied=False,
                         n_experts=4, n_experts_used=2),
}


def get_preset(name):
    return TINY[name] if name in TINY else PRESETS[name]


def synth_model(preset_name, mix, max_seq_len, seed=1234, sigma=0.02, max_batch=1, n_layers=None, vocab=None):
    """(arch, desc, tensors); tensors = {gguf_name: (ggml_type, ne, ndarray)}."""
    p = dict(get_preset(preset_name))
    if n_layers is not None:
        p["n_layers"] = n_layers
    if vocab is not None:
        p["vocab"] = vocab
    rng = np.random.default_rng(seed)
    desc = make_desc(p, max_seq_len, max_batch)
    tensors = {}
    for name, ttype, ne in tensor_plan(p, mix):
        n = int(np.prod(ne))
        if name.endswith("norm.weight"):
            data = (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        elif name.endswith(".bias"):
            data = (0.02 * rng.standard_normal(n)).astype(np.float32)
        elif ttype == F32:
            data = (sigma * rng.standard_normal(n)).astype(np.float32)
        else:
            s = sigma
            if name in ("token_embd.weight",):
                s = 1.0 if p.get("tied") is False else sigma * 4  # embeddings: O(1) rows (tied ones stay small for logits)
            w = (s * rng.standard_normal(n)).astype(np.float32)
            data = O.quantize(ttype, w)
        tensors[name] = (ttype, ne, data)
    return p["arch"], desc, tensors


def prompt_tokens(n, vocab):
    """prompt = (i*7919+1) % vocab"""
    return [(i * 7919 + 1) % vocab for i in range(n)]


def rel_err(a, b):
    """max |a-b| / max |b|: the relative error the 1e-3 parity bound is stated on."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))
