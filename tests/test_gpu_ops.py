"""Parity of the Backend per-op surface (through the C ABI) against the oracle.

Tolerances follow the reference's own GPU-vs-CPU tests (tests/metal_integration.rs:93-257:
add/mul 1e-5, silu/gelu 1e-4, rms_norm 1e-3, rope 1e-3, vec_mat 1e-2) and its fused-vs-dequant
recipe (examples/compare_backend_manual.rs:152-215), tightened where this backend can do
better: dequantisation and add/mul/scale are BIT-EXACT, everything else <= 1e-5 relative
(max|a-b| / max|b|)."""
import os

import numpy as np
import pytest

from synth import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dequant_golden.npz")
QTYPES = [2, 6, 8, 12, 13, 14]


@pytest.fixture(scope="module")
def be(b200):
    backend = b200.CudaB200Backend()
    assert backend.is_available(), "pytest -m gpu needs a CUDA device"
    return backend


@pytest.mark.parametrize("t", QTYPES)
def test_dequantize_golden_bit_exact(be, t):
    g = np.load(GOLD)
    raw, want = g[f"raw_{t}"], g[f"deq_{t}"]
    got = be.dequantize(raw.reshape(-1), t, want.size).reshape(want.shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("t", QTYPES + [0, 1])
def test_dequantize_random_weights_bit_exact_vs_oracle(be, oracle, t):
    rng = np.random.default_rng(100 + t)
    x = rng.normal(0, 0.02, 256 * 200).astype(np.float32)
    raw = oracle.quantize(t, x)
    want = oracle.dequantize(t, raw, x.size)
    got = be.dequantize(raw, t, x.size)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def gemv_err(got, want, k, w_sigma, x):
    """max |a-b| over the typical magnitude of an output (sigma_w * |x|_2): a single output can cancel to ~0 (n = 1),
    which would make max|b| a meaningless denominator.  Never smaller than the plain rel_err denominator."""
    scale = max(float(np.max(np.abs(want))), w_sigma * float(np.linalg.norm(x)))
    return float(np.max(np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)))) / scale


# (k, n) shapes: ragged row counts (n % 4 != 0), block counts that do not fill a warp step, the 8B shapes scaled down
GEMV_SHAPES = [(256, 1), (256, 7), (512, 33), (1024, 130), (4096, 64), (5632, 24), (14336, 9), (2048, 1000)]


@pytest.mark.parametrize("t", [12, 13, 14])
@pytest.mark.parametrize("k,n", GEMV_SHAPES)
def test_vec_mat_q_kquants(be, oracle, t, k, n):
    rng = np.random.default_rng(k * 131 + n + t)
    w = oracle.quantize(t, rng.normal(0, 0.05, k * n).astype(np.float32))
    x = rng.standard_normal(k).astype(np.float32)
    want = oracle.vec_mat_q(t, w, x, n)
    got = be.vec_mat_q(x, w, t, k, n)
    assert gemv_err(got, want, k, 0.05, x) < 1e-5


@pytest.mark.parametrize("t", [2, 6, 8, 0, 1])
@pytest.mark.parametrize("k,n", [(32, 3), (64, 17), (896, 130), (448, 64), (4096, 37), (4864, 12)])
def test_vec_mat_q_block32_types(be, oracle, t, k, n):
    rng = np.random.default_rng(k * 7 + n + t)
    w = oracle.quantize(t, rng.normal(0, 0.05, k * n).astype(np.float32))
    x = rng.standard_normal(k).astype(np.float32)
    want = oracle.vec_mat_q(t, w, x, n)
    got = be.vec_mat_q(x, w, t, k, n)
    assert gemv_err(got, want, k, 0.05, x) < 1e-5


def test_vec_mat_q_random_byte_blocks(be, oracle):
    """Arbitrary (non-quantiser) blocks incl. 6-bit scale packing of groups 4..7 and negative Q6_K scales."""
    g = np.load(GOLD)
    for t in [12, 13, 14, 8]:
        raw = g[f"raw_{t}"]
        nb_row = 4
        n = raw.shape[0] // nb_row
        k = nb_row * (256 if t >= 12 else 32)
        if k % 32:
            continue
        x = np.random.default_rng(t).standard_normal(k).astype(np.float32)
        want = oracle.vec_mat_q(t, raw.reshape(-1), x, n)
        got = be.vec_mat_q(x, raw.reshape(-1), t, k, n)
        assert rel_err(got, want) < 1e-5


def test_vec_mat_gguf_layout_kat(be):  # cpu/ops.rs:1813-1867
    w = np.array([1, 2, 3, 4, 5, 6], np.float32).reshape(2, 3)
    assert be.vec_mat(np.ones(3, np.float32), w).tolist() == [6.0, 15.0]
    w = np.array([1, 0, 0, 0, 1, 0], np.float32).reshape(2, 3)
    assert be.vec_mat(np.array([7, 8, 9], np.float32), w).tolist() == [7.0, 8.0]


def test_elementwise(be, oracle):
    rng = np.random.default_rng(1)
    a = rng.standard_normal(5000).astype(np.float32)
    b = rng.standard_normal(5000).astype(np.float32)
    assert np.array_equal(be.add(a, b), oracle.add(a, b))
    assert np.array_equal(be.mul(a, b), oracle.mul(a, b))
    assert np.array_equal(be.scale(a, 2.5), oracle.scale(a, 2.5))
    assert be.add(np.array([1, 2, 3, 4], np.float32), np.array([10, 20, 30, 40], np.float32)).tolist() == [11, 22, 33, 44]
    assert rel_err(be.silu(a), oracle.silu(a)) < 1e-6
    x = np.array([0.0, 1.0, -1.0, 2.0], np.float32)
    s = be.silu(x)
    assert abs(s[0]) < 1e-6 and abs(s[1] - 0.731) < 0.01 and abs(s[2] + 0.269) < 0.01
    want = 0.5 * a * (1.0 + np.tanh(0.7978846 * (a + 0.044715 * a ** 3)))
    assert rel_err(be.gelu(a), want) < 1e-5


def test_softmax(be, oracle):
    r = be.softmax(np.array([1, 2, 3, 4], np.float32))
    assert abs(r.sum() - 1.0) < 1e-6 and r[0] < r[1] < r[2] < r[3]
    x = np.random.default_rng(2).standard_normal(3001).astype(np.float32) * 4
    assert rel_err(be.softmax(x), oracle.softmax(x)) < 1e-5


@pytest.mark.parametrize("n", [4, 896, 4096, 8192])
def test_rms_norm(be, oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    assert rel_err(be.rms_norm(x, w, 1e-5), oracle.rms_norm(x, w, 1e-5)) < 1e-5
    r = be.rms_norm(np.array([1, 2, 3, 4], np.float32), np.ones(4, np.float32), 1e-5)
    assert abs(r[0] - 0.365) < 0.01 and abs(r[3] - 1.46) < 0.01


@pytest.mark.parametrize("neox", [False, True])
@pytest.mark.parametrize("pos", [0, 1, 77, 8191])
@pytest.mark.parametrize("hd,base", [(64, 1e4), (128, 5e5), (64, 1e6)])
def test_rope(be, oracle, neox, pos, hd, base):
    rng = np.random.default_rng(pos + hd)
    q = rng.standard_normal((8, 1, hd)).astype(np.float32)
    k = rng.standard_normal((2, 1, hd)).astype(np.float32)
    gq, gk = be.rope(q, k, pos, base, 1.0, neox)
    wq, wk = oracle.rope(q, k, pos, base, 1.0, neox)
    assert rel_err(gq, wq) < 2e-6 and rel_err(gk, wk) < 2e-6


def test_rope_kats(be):  # cpu/ops.rs:1688-1777
    q = np.array([1, 2, 3, 4], np.float32).reshape(1, 1, 4)
    r, _ = be.rope(q, q.copy(), 1, 10000.0, 1.0, True)
    r = r.ravel()
    assert abs(r[0] + 1.98) < 0.05 and abs(r[2] - 2.46) < 0.05 and abs(r[1] - 1.96) < 0.05 and abs(r[3] - 4.02) < 0.05
    q = np.array([1, 0, 0, 0], np.float32).reshape(1, 1, 4)
    r, _ = be.rope(q, q.copy(), 1, 10000.0, 1.0, False)
    assert abs(r.ravel()[0] - 0.54) < 0.02
    a, _ = be.rope(q, q.copy(), 4, 10000.0, 4.0, False)
    assert np.array_equal(a, r)


@pytest.mark.parametrize("nh,nkv,hd", [(8, 2, 128), (7, 1, 64), (8, 1, 64), (4, 4, 128), (32, 8, 128)])
@pytest.mark.parametrize("kv_len", [1, 2, 31, 257, 2048])
def test_attention_cached(be, oracle, nh, nkv, hd, kv_len):
    rng = np.random.default_rng(kv_len * 3 + nh)
    max_seq = kv_len + 5
    q = rng.standard_normal((nh, 1, hd)).astype(np.float32)
    kc = rng.standard_normal((nkv, max_seq, hd)).astype(np.float32)
    vc = rng.standard_normal((nkv, max_seq, hd)).astype(np.float32)
    scale = 1.0 / np.sqrt(hd)
    got = be.attention_cached(q, kc, vc, scale, kv_len)
    want = oracle.attention_cached(q, kc, vc, scale, kv_len)
    assert rel_err(got, want) < 1e-5


def test_attention_causal_kat(be):  # cpu/ops.rs:1780-1810
    q = np.array([1, 0, 0, 0, 0, 1, 0, 0], np.float32).reshape(1, 2, 4)
    v = np.arange(1, 9, dtype=np.float32).reshape(1, 2, 4)
    out = be.attention(q, q.copy(), v, 1.0 / np.sqrt(2.0))
    assert abs(out.ravel()[0] - 1.0) < 0.1
    out = be.attention(np.ones((4, 1, 4), np.float32), np.ones((2, 1, 4), np.float32), np.ones((2, 1, 4), np.float32), 0.5)
    assert np.all(np.isfinite(out))


def test_error_mapping(be, b200):
    with pytest.raises(b200.ShapeMismatch):
        be.add(np.ones(4, np.float32), np.ones(5, np.float32))
    with pytest.raises(b200.DTypeMismatch):
        be.add(np.ones(4, np.float64), np.ones(4, np.float64))
    with pytest.raises(b200.UnsupportedDType):
        be.dequantize(np.zeros(84, np.uint8), 10, 256)  # Q2_K is not on the path
    with pytest.raises(b200.ShapeMismatch):
        be.vec_mat_q(np.ones(256, np.float32), np.zeros(100, np.uint8), 12, 256, 1)
    with pytest.raises(b200.InvalidArgument):
        be.rope(np.ones((2, 4), np.float32), np.ones((2, 4), np.float32), 0, 1e4, 1.0, False)


# ---- tcgen05 / TMEM dequant-GEMM (csrc/gemm_umma.cuh): T rows of vec_mat_q at once, fp16 operands, f32 accumulation ----
# tolerance: 1e-3 of the typical output magnitude (fp16 rounding of both operands, ~3e-4 measured), stated here because
# this path trades the exact f32 arithmetic of the GEMV for the tensor cores (prefill / batched decode).
@pytest.mark.parametrize("t", [12, 13, 14, 8])
@pytest.mark.parametrize("rows,k,n", [(1, 256, 5), (32, 512, 130), (33, 1024, 128), (70, 4096, 257), (300, 2048, 64)])
def test_mat_mat_q_tensor_core_gemm(be, oracle, t, rows, k, n):
    rng = np.random.default_rng(k * 17 + n + t + rows)
    w = oracle.quantize(t, rng.normal(0, 0.05, k * n).astype(np.float32))
    x = rng.standard_normal((rows, k)).astype(np.float32)
    got = be.mat_mat_q(x, w, t, k, n)
    assert got.shape == (rows, n)
    for r in sorted({0, rows // 2, rows - 1}):
        want = oracle.vec_mat_q(t, w, x[r], n)
        assert gemv_err(got[r], want, k, 0.05, x[r]) < 1e-3, f"row {r}"
