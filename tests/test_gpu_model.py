"""End-to-end parity of GpuOnlyInference (C ABI) against the oracle's LlamaModel::forward on
the same synthetic random-init weights: per-layer hidden states and logits within 1e-3
relative (max|a-b| / max|b|, the north-star tolerance), greedy token sequences identical."""
import os

import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def build_pair(b200, oracle, preset, mix, max_seq, taps=False, **kw):
    arch, desc, tensors = synth.synth_model(preset, mix, max_seq, **kw)
    gpu = b200.GpuOnlyInference(desc, tensors, taps=taps)
    ref = oracle.OracleModel(desc, tensors)
    return desc, gpu, ref


CASES = [("llama-tiny", "Q4_K_M"), ("llama-tiny", "Q5_K_M"), ("llama-tiny", "Q8_0"), ("llama-tiny", "Q6_K"),
         ("qwen-tiny", "Q4_K_M"), ("tinyllama-tiny", "Q8_0"), ("tinyllama-tiny", "Q6_K"), ("tinyllama-tiny", "Q4_0"),
         ("mixtral-tiny", "Q5_K_M"), ("mixtral-tiny", "Q4_K_M")]


@pytest.mark.parametrize("preset,mix", CASES)
def test_hidden_states_and_logits_per_layer(b200, oracle, preset, mix):
    desc, gpu, ref = build_pair(b200, oracle, preset, mix, 64, taps=True)
    toks = synth.prompt_tokens(12, desc["vocab"])
    for i, t in enumerate(toks):
        want = ref.forward([t])
        got = gpu.forward(t)
        for layer in range(desc["n_layers"] + 1):
            e = rel_err(gpu.hidden(layer), ref.hidden(layer))
            assert e < TOL, f"token {i} layer {layer}: rel err {e}"
        assert rel_err(got, want) < TOL, f"token {i} logits"
    assert gpu.position() == len(toks) == ref.position()
    gpu.close()


@pytest.mark.parametrize("preset,mix", CASES)
def test_greedy_tokens_identical(b200, oracle, preset, mix):
    """bench protocol (src/main.rs:1796-1834): prefill, then greedy decode with raw-logit argmax."""
    desc, gpu, ref = build_pair(b200, oracle, preset, mix, 96)
    prompt = synth.prompt_tokens(16, desc["vocab"])
    want_logits = ref.forward(prompt)
    got_logits = b200.GpuModelWrapper(gpu).forward(prompt, 0)
    assert rel_err(got_logits, want_logits) < TOL
    tok_r, tok_g = oracle.argmax_last(want_logits), oracle.argmax_last(got_logits)
    seq_r, seq_g = [], []
    for _ in range(48):
        seq_r.append(tok_r)
        seq_g.append(tok_g)
        tok_r = oracle.argmax_last(ref.forward([tok_r]))
        tok_g = oracle.argmax_last(gpu.forward(tok_g))
    assert seq_g == seq_r
    gpu.close()


def test_device_greedy_matches_host_greedy(b200, oracle):
    """b200_decode_greedy (argmax on the device, graph replay) == forward + host argmax == oracle."""
    desc, gpu, ref = build_pair(b200, oracle, "llama-tiny", "Q4_K_M", 96)
    prompt = synth.prompt_tokens(8, desc["vocab"])
    first = oracle.argmax_last(b200.GpuModelWrapper(gpu).forward(prompt, 0))
    toks, ms = gpu.decode_greedy(first, 40)
    want = []
    tok = oracle.argmax_last(ref.forward(prompt))
    assert tok == first
    for _ in range(40):
        tok = oracle.argmax_last(ref.forward([tok]))
        want.append(tok)
    assert toks.tolist() == want and ms > 0
    assert gpu.position() == 8 + 40
    gpu.close()


def test_reset_and_wrapper_semantics(b200, oracle):
    """GpuModelWrapper resets the engine when ctx.position == 0 (backend/mod.rs:334-336)."""
    desc, gpu, ref = build_pair(b200, oracle, "qwen-tiny", "Q4_K_M", 64)
    w = b200.GpuModelWrapper(gpu)
    p1 = synth.prompt_tokens(6, desc["vocab"])
    a = w.forward(p1, 0)
    assert gpu.position() == 6
    b = w.forward(p1, 0)  # new sequence: must reset and reproduce
    assert gpu.position() == 6 and np.array_equal(a, b)
    gpu.reset()
    assert gpu.position() == 0
    with pytest.raises(b200.InvalidArgument):
        w.forward([], 0)
    with pytest.raises(b200.InvalidArgument):
        gpu.forward(desc["vocab"])  # token id out of range (llama.rs:297-303)
    gpu.close()


def test_context_length_exceeded(b200, oracle):
    desc, gpu, ref = build_pair(b200, oracle, "tinyllama-tiny", "Q8_0", 8)
    for t in range(8):
        gpu.prefill_token(t)
    with pytest.raises(b200.InvalidArgument):
        gpu.forward(1)
    gpu.close()


def test_independent_sequence_slots(b200, oracle):
    """b200_decode_batch: slots do not clobber each other's KV cache."""
    arch, desc, tensors = synth.synth_model("llama-tiny", "Q4_K_M", 48, max_batch=3)
    gpu = b200.GpuOnlyInference(desc, tensors)
    refs = [oracle.OracleModel(desc, tensors) for _ in range(3)]
    streams = [synth.prompt_tokens(10, desc["vocab"]), [5, 9, 2, 700, 31, 8, 8, 8, 1, 0], list(range(100, 110))]
    for step in range(10):
        toks = [s[step] for s in streams]
        got = gpu.decode_batch([0, 1, 2], toks)
        for i in range(3):
            want = refs[i].forward([toks[i]])
            assert rel_err(got[i], want) < TOL
    with pytest.raises(b200.InvalidArgument):
        gpu.decode_batch([0, 0], [1, 2])
    gpu.close()


def test_loads_from_a_gguf_file(b200, oracle, tmp_path):
    """Synthetic GGUF written with the metadata types the reference loader demands, read back through
    the host shim (gguf_io mirrors ModelLoader::parse_config) — same logits as the in-memory model."""
    from llama_gguf_b200 import gguf_io

    arch, desc, tensors = synth.synth_model("qwen-tiny", "Q4_K_M", 32)
    path = os.path.join(tmp_path, "qwen-tiny.gguf")
    gguf_io.write_gguf(path, arch, desc, tensors)
    arch2, desc2, tensors2 = gguf_io.load_gguf(path)
    assert arch2 == "qwen2" and desc2["rope_neox"] == 1 and desc2["tied_output"] == 1
    gpu = b200.GpuOnlyInference.from_model((desc2, tensors2), 32)
    ref = oracle.OracleModel(desc, tensors)
    toks = synth.prompt_tokens(5, desc["vocab"])
    assert rel_err(gpu.forward_batch(toks), ref.forward(toks)) < TOL
    gpu.close()


def test_upload_validation(b200):
    arch, desc, tensors = synth.synth_model("tinyllama-tiny", "Q8_0", 16)
    bad = dict(tensors)
    del bad["blk.1.ffn_down.weight"]
    with pytest.raises(b200.InvalidArgument):
        b200.GpuOnlyInference(desc, bad)
    bad = dict(tensors)
    t, ne, data = bad["blk.0.attn_q.weight"]
    bad["blk.0.attn_q.weight"] = (t, [ne[0], ne[1] // 2], data[: data.size // 2])
    with pytest.raises(b200.ShapeMismatch):
        b200.GpuOnlyInference(desc, bad)
    with pytest.raises(b200.InvalidArgument):
        b200.GpuOnlyInference(dict(desc, n_kv_heads=3), tensors)  # n_heads % n_kv_heads != 0
