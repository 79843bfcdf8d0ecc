"""INT8 KV cache: the reference's QuantizedKVCache in its Int8 format (src/model/kv_quantized.rs).

CPU part: the oracle's restatement of quantize_int8 / dequantize_int8 against the reference's own tests
(kv_quantized.rs:547-561 test_int8_roundtrip, :613-670 test_quantized_kv_cache_basic — same inputs, same bounds) and against an
independent numpy statement of the formula (bit-exact).  GPU part: a model decoded with the int8 cache (csrc/kv_int8.cuh:
quantised write, attention over q * scale) against the oracle model whose cache rows pass through the same format."""
import numpy as np
import pytest

import synth
from synth import rel_err


def _np_quantize_int8(x):
    x = np.asarray(x, dtype=np.float32)
    max_abs = np.float32(np.max(np.abs(x))) if x.size else np.float32(0)
    scale = np.float32(max_abs / np.float32(127.0)) if max_abs > np.float32(1e-10) else np.float32(1.0)
    r = (x / scale).astype(np.float32)
    q = np.where(r >= 0, np.floor(r + np.float32(0.5)), np.ceil(r - np.float32(0.5)))   # f32::round: half away from zero
    # floor(r + 0.5) can differ from round-half-away when r + 0.5 is not representable; exact form:
    q = np.sign(r) * np.floor(np.abs(r).astype(np.float64) + 0.5)
    return np.clip(q, -128, 127).astype(np.int8), float(scale)


def test_reference_int8_roundtrip_kat(oracle):
    data = (np.arange(128, dtype=np.float32) * np.float32(0.1) - np.float32(6.4)).astype(np.float32)   # kv_quantized.rs:549
    q, s = oracle.kv_quantize_int8(data)
    dec = oracle.kv_dequantize_int8(q, s)
    for o, d in zip(data, dec):
        err = abs(o - d) / abs(o) if abs(o) > 1e-6 else abs(o - d)
        assert err < 0.02                                                                               # kv_quantized.rs:559


def test_reference_quantized_kv_cache_basic_kat(oracle):
    nkv, hd = 4, 64                                                                                     # kv_quantized.rs:615-618
    k = (np.arange(nkv * hd, dtype=np.float32) * np.float32(0.01) - np.float32(1.0)).astype(np.float32)
    v = (np.arange(nkv * hd, dtype=np.float32) * np.float32(0.02) - np.float32(0.5)).astype(np.float32)
    for row in (k[:hd], v[:hd]):                                                                        # head 0, position 0
        q, s = oracle.kv_quantize_int8(row)
        back = oracle.kv_dequantize_int8(q, s)
        for a, b in zip(row, back):
            err = abs(a - b) / abs(a) if abs(a) > 1e-6 else abs(a - b)
            assert err < 0.15                                                                           # kv_quantized.rs:647


@pytest.mark.parametrize("seed", range(6))
def test_quantize_int8_matches_the_formula_bit_for_bit(oracle, seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(128) * 10.0 ** rng.integers(-6, 3)).astype(np.float32)
    if seed == 0:
        x[:] = 0.0                       # all-zero row: scale 1.0 (kv_quantized.rs:372-376)
    if seed == 1:
        x[5] = -x.max() * 2              # the largest magnitude is negative: q = -127, never -128
    q, s = oracle.kv_quantize_int8(x)
    q2, s2 = _np_quantize_int8(x)
    assert s == s2 and np.array_equal(q, q2)
    assert np.array_equal(oracle.kv_dequantize_int8(q, s), q.astype(np.float32) * np.float32(s))
    if seed:
        assert np.max(np.abs(q)) == 127


@pytest.mark.gpu
@pytest.mark.parametrize("preset,mix", [("llama-tiny", "Q4_K_M"), ("tinyllama-tiny", "Q8_0"), ("qwen-tiny", "Q4_K_M"), ("mixtral-tiny", "Q4_K_M")])
def test_int8_kv_decode_matches_the_oracle(b200, oracle, preset, mix):
    """Logits within 1e-3 of the oracle and identical greedy tokens, long enough (70 positions) that several attention splits hold
    rows; hd = 128 and 64, G = 4 / 8 / 7, NeoX RoPE with biases, MoE."""
    arch, desc, tensors = synth.synth_model(preset, mix, 96)
    gpu = b200.GpuOnlyInference(desc, tensors, kv_format="int8")
    assert gpu.kv_format() == "int8" and gpu.path() == "graph"
    assert gpu.stats()["kv_bytes_per_pos"] == 2 * desc["n_layers"] * desc["n_kv_heads"] * (desc["head_dim"] + 4)
    ref = oracle.OracleModel(desc, tensors, kv_format="int8")
    exact = oracle.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(40, desc["vocab"])
    want = ref.forward(prompt)
    got = gpu.prefill(prompt)
    assert rel_err(got, want) < 1e-3
    assert rel_err(want, exact.forward(prompt)) > 1e-5, "the int8 cache must actually change the logits (else the test proves nothing)"
    tok = oracle.argmax_last(want)
    for _ in range(30):
        want = ref.forward([tok])
        got = gpu.forward(tok)
        assert rel_err(got, want) < 1e-3
        assert oracle.argmax_last(got) == oracle.argmax_last(want)
        tok = oracle.argmax_last(want)
    assert gpu.position() == 70 == ref.position()
    gpu.reset()
    ref.reset()
    assert rel_err(gpu.prefill(prompt[:9]), ref.forward(prompt[:9])) < 1e-3
    gpu.close()


@pytest.mark.gpu
def test_int8_kv_rejects_what_it_does_not_cover(b200):
    arch, desc, tensors = synth.synth_model("llama-tiny", "Q4_K_M", 64)
    with pytest.raises(b200.Unsupported):
        b200.GpuOnlyInference(desc, tensors, kv_format="fp8")


@pytest.mark.gpu
@pytest.mark.parametrize("preset,mix,n", [("llama-stream-tiny", "Q4_K_M", 130), ("tinyllama-stream-tiny", "Q8_0", 70), ("llama-tiny", "Q4_K_M", 48)])
def test_int8_kv_prompt_through_the_tensor_core_path_is_opt_in(b200, oracle, preset, mix, n, monkeypatch):
    """OPT-IN (B200_KV_INT8_GEMM=1), and NOT inside north_star's 1e-3: a prompt of >= 32 tokens takes ONE pass of the dequant-GEMMs
    -- RoPE + quantised cache write for all rows (prefill_rope_kv_q8_kernel), tcgen05 attention over fp16 tiles of the DEQUANTISED
    rows (prefill_kv16_q8_kernel) -- but its K / V rows differ from the exact path's by the fp16 operand rounding before they are
    quantised, and rounding to int8 codes is discontinuous (the neighbouring code is 1/127 of the row maximum away).  Measured
    8e-4 .. 7e-3 against the oracle; held to 1e-2 here, which is why the default prompt path of an int8 context stays the exact
    token-by-token one (held to 1e-3 by test_int8_kv_decode_matches_the_oracle: its 40-token prompt must NOT take this path)."""
    arch, desc, tensors = synth.synth_model(preset, mix, 256)
    ref = oracle.OracleModel(desc, tensors, kv_format="int8")
    prompt = synth.prompt_tokens(n, desc["vocab"])
    want0 = ref.forward(prompt)
    # default: exact path, one launch sequence per token, 1e-3
    gpu = b200.GpuOnlyInference(desc, tensors, kv_format="int8")
    l0 = gpu.stats()["kernel_launches"]
    assert rel_err(gpu.prefill(prompt), want0) < 1e-3
    assert gpu.stats()["kernel_launches"] - l0 > 10 * n
    gpu.close()
    monkeypatch.setenv("B200_KV_INT8_GEMM", "1")
    gpu = b200.GpuOnlyInference(desc, tensors, kv_format="int8")
    monkeypatch.delenv("B200_KV_INT8_GEMM")
    l0 = gpu.stats()["kernel_launches"]
    got = gpu.prefill(prompt)
    assert gpu.stats()["kernel_launches"] - l0 < 40 * desc["n_layers"] + 20, "one GEMM pass, not one launch sequence per token"
    e0 = rel_err(got, want0)
    tok = oracle.argmax_last(want0)
    worst = e0
    for _ in range(6):                                   # per-op decode over the rows the GEMM pass quantised (teacher-forced)
        want = ref.forward([tok])
        worst = max(worst, rel_err(gpu.forward(tok), want))
        tok = oracle.argmax_last(want)
    more = synth.prompt_tokens(40, desc["vocab"])[::-1]
    e2 = rel_err(gpu.prefill(more), ref.forward(more))   # appended chunk: attention over old int8 rows + its own
    print(f"int8 KV GEMM prompt {preset} n={n}: prompt {e0:.2e}, decode after it {worst:.2e}, appended prompt {e2:.2e}")
    assert max(e0, worst, e2) < 1e-2
    assert gpu.position() == n + 6 + 40 == ref.position()
    gpu.close()
