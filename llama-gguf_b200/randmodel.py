"""Random-init weights written DIRECTLY as finite GGUF blocks (numpy only).

Used by bench.py for the full-size models: pure-bandwidth work needs valid blocks with
sane magnitudes, not a quantiser.  f16 scale fields are set so that dequantised weights
have zero mean and standard deviation ~sigma (SURVEY.md §8d "direct random finite
blocks").  Parity tests do NOT use this module: they quantise N(0, sigma^2) weights with
the restated reference quantisers (tests/synth.py).
"""
import numpy as np

from .presets import F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K, make_desc, tensor_plan

# random_model(repeat_bytes=N): tensors above N bytes repeat their first N bytes of random blocks -- host generation time of the
# 40 GB presets (bench.py uses it for the 70B / 8x7B lines only; HBM traffic does not depend on the byte values).  None: off.

BLOCK = {F32: (1, 4), F16: (1, 2), Q4_0: (32, 18), Q5_0: (32, 22), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176), Q6_K: (256, 210)}


def _f16_bits(v):
    return int(np.array([v], dtype=np.float16).view(np.uint16)[0])


def random_blocks(ttype, n_elems, rng, sigma=0.02, repeat_bytes=None):
    """uint8 array holding n_elems/bs random blocks of `ttype` whose values are ~ (0, sigma^2)."""
    be, bb = BLOCK[ttype]
    nb = n_elems // be
    if ttype == F32:
        return rng.standard_normal(n_elems, dtype=np.float32) * np.float32(sigma)
    if ttype == F16:
        return (rng.standard_normal(n_elems, dtype=np.float32) * np.float32(sigma)).astype(np.float16)
    if repeat_bytes and nb * bb > repeat_bytes:
        nb0 = max(1, repeat_bytes // bb)
        raw = np.resize(rng.integers(0, 256, size=(nb0, bb), dtype=np.uint8), (nb, bb))
    else:
        raw = rng.integers(0, 256, size=(nb, bb), dtype=np.uint8)
    u16 = raw.view(np.uint16)  # every block size is even
    if ttype == Q4_K:   # d*sc*q - dmin*m, sc,m in [0,63], q in [0,15]: zero mean when dmin = 7.5 d
        d = sigma / 200.0
        u16[:, 0] = _f16_bits(d)
        u16[:, 1] = _f16_bits(7.5 * d)
    elif ttype == Q5_K:  # q in [0,31]
        d = sigma / 400.0
        u16[:, 0] = _f16_bits(d)
        u16[:, 1] = _f16_bits(15.5 * d)
    elif ttype == Q6_K:  # d*sc*(q-32), sc int8
        u16[:, 104] = _f16_bits(sigma / 1400.0)
    elif ttype == Q8_0:  # d*q, q int8
        u16[:, 0] = _f16_bits(sigma / 74.0)
    elif ttype == Q5_0:  # d*(q-16)
        u16[:, 0] = _f16_bits(sigma / 9.2)
    elif ttype == Q4_0:  # d*(q-8)
        u16[:, 0] = _f16_bits(sigma / 4.6)
    return raw.reshape(-1)


def random_model(preset, mix, max_seq_len, seed=1234, sigma=0.02, max_batch=1, upload=None, repeat_bytes=None):
    """(desc, tensors) for `preset`; if `upload(name, type, ne, data)` is given tensors are streamed to it
    one at a time instead of being kept (the 8B model is 4.6 GB of host memory otherwise)."""
    rng = np.random.default_rng(seed)
    desc = make_desc(preset, max_seq_len, max_batch)
    tensors = {}
    for name, ttype, ne in tensor_plan(preset, mix):
        n = int(np.prod(ne))
        if name.endswith("norm.weight"):
            data = (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        elif name.endswith(".bias"):
            data = (sigma * rng.standard_normal(n)).astype(np.float32)
        else:
            data = random_blocks(ttype, n, rng, sigma, repeat_bytes)
        if upload is not None:
            upload(name, ttype, ne, data)
        else:
            tensors[name] = (ttype, ne, data)
    return desc, tensors
