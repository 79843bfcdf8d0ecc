"""Pins the oracle's block dequantisation bit-for-bit against the gguf-py golden vectors
(tests/golden/make_golden.py) — the only byte->float authority available for K-quants,
since the reference's own tests hold no exact K-quant vectors."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dequant_golden.npz")
TYPES = [2, 6, 8, 12, 13, 14]


@pytest.mark.parametrize("t", TYPES)
def test_oracle_dequant_matches_golden_bit_exact(oracle, t):
    g = np.load(GOLD)
    raw, want = g[f"raw_{t}"], g[f"deq_{t}"]
    got = oracle.dequantize(t, raw.reshape(-1), want.size).reshape(want.shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("t", TYPES)
def test_oracle_fused_dot_close_to_dequant_then_dot(oracle, t):
    """examples/compare_backend_manual.rs:169 — fused vec_mat_q vs dequantise-then-dot, abs tol 1e-3."""
    g = np.load(GOLD)
    raw, deq = g[f"raw_{t}"], g[f"deq_{t}"]
    rng = np.random.default_rng(t)
    k = deq.shape[1] * 4  # 4 blocks per row, 6 rows
    w = raw.reshape(6, -1)
    x = rng.standard_normal(k).astype(np.float32)
    got = oracle.vec_mat_q(t, w, x, 6)
    want = deq.reshape(6, k).astype(np.float64) @ x.astype(np.float64)
    assert np.max(np.abs(got - want)) < 1e-3 * max(1.0, np.max(np.abs(want)))


def test_f16_conversion_exhaustive(oracle):
    """half::f16::to_f32 is exact for all 65536 bit patterns; from_f32 round-trips them."""
    bits = np.arange(65536, dtype=np.uint16)
    want = bits.view(np.float16).astype(np.float32)
    got = np.array([oracle.lib().orc_f16_to_f32(int(b)) for b in bits[::7]], dtype=np.float32)
    w = want[::7]
    ok = (got.view(np.uint32) == w.view(np.uint32)) | (np.isnan(got) & np.isnan(w))
    assert ok.all()
    finite = np.isfinite(w)
    back = np.array([oracle.lib().orc_f32_to_f16(float(v)) for v in w[finite]], dtype=np.uint16)
    assert np.array_equal(back, bits[::7][finite])


def test_f32_to_f16_rounding_matches_numpy(oracle):
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.standard_normal(4000).astype(np.float32) * 10.0 ** rng.integers(-9, 5, 4000),
                        np.array([65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 0.0, -0.0], dtype=np.float32)])
    want = v.astype(np.float16).view(np.uint16)
    got = np.array([oracle.lib().orc_f32_to_f16(float(x)) for x in v], dtype=np.uint16)
    assert np.array_equal(got, want)
