"""Generates tests/golden/dequant_golden.npz with gguf-py (gguf==0.19.0), an implementation
independent of both the reference and this repo whose numpy dequantize uses the same
operation order as the reference's dequant.rs.  For every block format on
the hot path: 24 random blocks (random bytes, finite f16 scales incl. one subnormal and
one negative scale) and their dequantised f32 values.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os

import numpy as np
from gguf import GGMLQuantizationType as GT
from gguf import quants

TYPES = {2: GT.Q4_0, 6: GT.Q5_0, 8: GT.Q8_0, 12: GT.Q4_K, 13: GT.Q5_K, 14: GT.Q6_K}
BLOCK = {2: (32, 18), 6: (32, 22), 8: (32, 34), 12: (256, 144), 13: (256, 176), 14: (256, 210)}
SCALE_OFFS = {2: [0], 6: [0], 8: [0], 12: [0, 2], 13: [0, 2], 14: [208]}


def main():
    rng = np.random.default_rng(20240607)
    out = {}
    for t, gt in TYPES.items():
        be, bb = BLOCK[t]
        nb = 24
        raw = rng.integers(0, 256, size=(nb, bb), dtype=np.uint8)
        for off in SCALE_OFFS[t]:
            sc = rng.normal(0, 0.05, nb).astype(np.float16)
            sc[0] = np.float16(6e-8)   # subnormal half
            sc[1] = np.float16(-0.031)
            sc[2] = np.float16(0.0)
            raw[:, off:off + 2] = sc.view(np.uint8).reshape(nb, 2)
        deq = quants.dequantize(raw, gt).astype(np.float32)
        assert np.isfinite(deq).all()
        out[f"raw_{t}"] = raw
        out[f"deq_{t}"] = deq.reshape(nb, be)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dequant_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
