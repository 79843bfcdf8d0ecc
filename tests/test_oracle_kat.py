"""The reference's own known-answer tests for this path, restated against the oracle
(SURVEY.md §4 rows 1-2, §8c).  Each test cites the reference test it restates."""
import numpy as np
import pytest


def test_add_mul_scale_exact(oracle):  # cpu/ops.rs:1568-1600
    a = np.array([1, 2, 3, 4], np.float32)
    assert oracle.add(a, np.array([10, 20, 30, 40], np.float32)).tolist() == [11, 22, 33, 44]
    assert oracle.mul(a, np.array([2, 3, 4, 5], np.float32)).tolist() == [2, 6, 12, 20]
    assert oracle.scale(a, 2.5).tolist() == [2.5, 5.0, 7.5, 10.0]


def test_silu(oracle):  # cpu/ops.rs:1602-1617
    r = oracle.silu(np.array([0, 1, -1, 2], np.float32))
    assert abs(r[0]) < 1e-6 and abs(r[1] - 0.731) < 0.01 and abs(r[2] + 0.269) < 0.01


def test_softmax(oracle):  # cpu/ops.rs:1619-1633
    r = oracle.softmax(np.array([1, 2, 3, 4], np.float32))
    assert abs(r.sum() - 1.0) < 1e-6 and r[0] < r[1] < r[2] < r[3]


def test_rms_norm(oracle):  # cpu/ops.rs:1636-1648 and cpu/simd.rs:1200-1220
    x = np.array([1, 2, 3, 4], np.float32)
    r = oracle.rms_norm(x, np.ones(4, np.float32), 1e-5)
    assert abs(r[0] - 0.365) < 0.01 and abs(r[3] - 1.46) < 0.01
    r = oracle.rms_norm(x, np.ones(4, np.float32), 1e-6)
    assert np.max(np.abs(r - x / np.sqrt(30.0 / 4.0))) < 1e-5


def test_dot_and_axpy(oracle):  # cpu/simd.rs:1176-1183
    a = np.arange(1, 9, dtype=np.float32)
    assert abs(oracle.dot_f32(a, np.ones(8, np.float32)) - 36.0) < 1e-6


def test_silu_mul_inplace(oracle):  # cpu/simd.rs:1235-1253
    gate = np.array([1.0, -1.0, 2.0, 0.0, 0.5, -0.5, 3.0, -2.0], np.float32)
    up = np.array([2.0, 3.0, 1.0, 5.0, 4.0, 2.0, 0.5, 1.0], np.float32)
    want = gate / (1.0 + np.exp(-gate)) * up
    assert np.max(np.abs(oracle.silu_mul(gate, up) - want)) < 1e-5


def test_rope_position_zero_is_identity(oracle):  # cpu/ops.rs:1688-1706
    q = np.array([1, 0, 1, 0, 0, 1, 0, 1], np.float32).reshape(2, 1, 4)
    rq, rk = oracle.rope(q, q.copy(), 0, 10000.0, 1.0, True)
    assert np.max(np.abs(rq - q)) < 1e-5 and np.max(np.abs(rk - q)) < 1e-5


def test_rope_normal_pairing(oracle):  # cpu/ops.rs:1708-1727
    q = np.array([1, 0, 0, 0], np.float32).reshape(1, 1, 4)
    rq, _ = oracle.rope(q, q.copy(), 1, 10000.0, 1.0, False)
    assert abs(rq.ravel()[0] - 0.54) < 0.02


def test_rope_neox_pairing(oracle):  # cpu/ops.rs:1729-1777
    q = np.array([1, 2, 3, 4], np.float32).reshape(1, 1, 4)
    r, _ = oracle.rope(q, q.copy(), 1, 10000.0, 1.0, True)
    r = r.ravel()
    assert abs(r[0] + 1.98) < 0.05 and abs(r[2] - 2.46) < 0.05 and abs(r[1] - 1.96) < 0.05 and abs(r[3] - 4.02) < 0.05


def test_rope_linear_scale_divides_position(oracle):  # cpu/ops.rs:1300
    q = np.array([1, 2, 3, 4], np.float32).reshape(1, 1, 4)
    a, _ = oracle.rope(q, q.copy(), 4, 10000.0, 4.0, False)
    b, _ = oracle.rope(q, q.copy(), 1, 10000.0, 1.0, False)
    assert np.array_equal(a, b)


def test_vec_mat_gguf_layout(oracle):  # cpu/ops.rs:1813-1867
    w = np.array([1, 2, 3, 4, 5, 6], np.float32)
    assert oracle.vec_mat_q(oracle.F32, w, np.ones(3, np.float32), 2).tolist() == [6.0, 15.0]
    w = np.array([1, 0, 0, 0, 1, 0], np.float32)
    assert oracle.vec_mat_q(oracle.F32, w, np.array([7, 8, 9], np.float32), 2).tolist() == [7.0, 8.0]


def test_attention_cached_gqa_uniform(oracle):  # cpu/ops.rs:1795-1810 (GQA 4:2, all-ones)
    q = np.ones((4, 1, 4), np.float32)
    kc = np.ones((2, 3, 4), np.float32)
    vc = np.ones((2, 3, 4), np.float32)
    out = oracle.attention_cached(q, kc, vc, 0.5, 3)
    assert np.all(np.isfinite(out)) and np.max(np.abs(out - 1.0)) < 1e-6


def test_attention_cached_first_position_returns_v0(oracle):  # cpu/ops.rs:1780-1793
    q = np.array([1, 0, 0, 0], np.float32).reshape(1, 1, 4)
    kc = np.zeros((1, 4, 4), np.float32)
    vc = np.zeros((1, 4, 4), np.float32)
    kc[0, 0] = [1, 0, 0, 0]
    vc[0, 0] = [1, 2, 3, 4]
    out = oracle.attention_cached(q, kc, vc, 1.0 / np.sqrt(2.0), 1)
    assert np.array_equal(out.ravel(), np.array([1, 2, 3, 4], np.float32))


def test_q4_0_q8_0_roundtrips(oracle):  # tests/dequant_test.rs:8-64, 66-88, 123-141
    x = ((np.arange(32) - 16) * 0.1).astype(np.float32)
    assert np.max(np.abs(oracle.dequantize(oracle.Q4_0, oracle.quantize(oracle.Q4_0, x), 32) - x)) < 0.15
    assert np.max(np.abs(oracle.dequantize(oracle.Q8_0, oracle.quantize(oracle.Q8_0, x), 32) - x)) < 0.02
    z = np.zeros(32, np.float32)
    assert np.all(oracle.dequantize(oracle.Q4_0, oracle.quantize(oracle.Q4_0, z), 32) == 0.0)
    assert np.all(oracle.dequantize(oracle.Q8_0, oracle.quantize(oracle.Q8_0, z), 32) == 0.0)
    x = ((np.arange(32) - 16) * 0.01).astype(np.float32)
    d = oracle.dequantize(oracle.Q8_0, oracle.quantize(oracle.Q8_0, x), 32)
    assert np.sqrt(np.mean((d - x) ** 2)) < 0.005


def test_q5_0_fixed_block(oracle):  # tests/dequant_test.rs:159-184: d=0.1, qh=0, qs=0x88 -> -0.8
    blk = np.zeros(22, np.uint8)
    blk[0:2] = np.array([0.1], np.float16).view(np.uint8)
    blk[6:] = 0x88
    d = oracle.dequantize(oracle.Q5_0, blk, 32)
    assert np.max(np.abs(d + 0.8)) < 0.01


def test_q8_0_symmetry(oracle):  # tests/dequant_test.rs:214-239
    p = (np.arange(32) * 0.1).astype(np.float32)
    a = oracle.dequantize(oracle.Q8_0, oracle.quantize(oracle.Q8_0, p), 32)
    b = oracle.dequantize(oracle.Q8_0, oracle.quantize(oracle.Q8_0, -p), 32)
    assert np.max(np.abs(a + b)) < 0.02


@pytest.mark.parametrize("t,bound", [(12, 4.5), (13, 4.5)])
def test_k_quant_roundtrip_rmse(oracle, t, bound):  # tensor/quant/dequant.rs:1226-1266
    x = ((np.arange(256) - 128) * 0.1).astype(np.float32)
    d = oracle.dequantize(t, oracle.quantize(t, x), 256)
    assert np.sqrt(np.mean((d - x) ** 2)) < bound


def test_q6_k_roundtrip(oracle):  # tensor/quant/dequant.rs:1268-1285
    x = ((np.arange(256) - 128) * 0.1).astype(np.float32)
    d = oracle.dequantize(oracle.Q6_K, oracle.quantize(oracle.Q6_K, x), 256)
    assert np.max(np.abs(d - x)) < 1.0


def test_moe_router_topk_softmax(oracle):  # model/moe.rs:505-517 (+ stable-sort tie rule :168)
    h = np.full(64, 0.1, np.float32)
    rng = np.random.default_rng(0)
    w = rng.standard_normal((4, 64)).astype(np.float32)
    idx, wts = oracle.moe_route(h, w, 4, 2)
    assert len(idx) == 2 and abs(wts.sum() - 1.0) < 0.01
    logits = w @ h
    assert idx[0] == int(np.argmax(logits))
    w[:] = 1.0  # all logits equal -> lowest indices win, equal weights
    idx, wts = oracle.moe_route(h, w, 4, 2)
    assert idx.tolist() == [0, 1] and abs(wts[0] - 0.5) < 1e-6


def test_argmax_last_wins(oracle):  # main.rs:1816-1821 (max_by keeps the last maximum)
    assert oracle.argmax_last(np.array([1, 5, 3, 5, 2], np.float32)) == 3
