"""The streamed megakernels (csrc/stream2.cuh, the decode path: TMA producer warp + mbarrier ring across phase boundaries,
14 consumer warps in pairs, stream-K dealing, loader-warp phase boundary; csrc/stream.cuh, its first generation) against the
oracle's LlamaModel::forward (src/model/llama.rs:275-362) and against the first megakernel, through the C ABI.
Tolerance: logits within 1e-3 relative (max|a-b| / max|b|), greedy token sequences identical (north star)."""
import os

import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3

CASES = [("llama-stream-tiny", "Q4_K_M"), ("llama-stream-tiny", "Q5_K_M"), ("llama-stream-tiny", "Q6_K"),
         ("llama-stream-tiny", "Q8_0"), ("llama-stream-tiny", "Q4_K"), ("tinyllama-stream-tiny", "Q4_K_M"),
         ("tinyllama-stream-tiny", "Q8_0")]


PATH = {2: "stream2", 1: "stream", 0: "mega"}


def _ctx(b200, desc, tensors, gen=2):
    """gen 2: stream2.cuh (default path), 1: stream.cuh, 0: mega.cuh"""
    os.environ["B200_STREAM"] = "1" if gen >= 1 else "0"
    os.environ["B200_STREAM2"] = "1" if gen >= 2 else "0"
    try:
        return b200.GpuOnlyInference(desc, tensors)
    finally:
        os.environ.pop("B200_STREAM", None)
        os.environ.pop("B200_STREAM2", None)


@pytest.mark.parametrize("gen", [2, 1])
@pytest.mark.parametrize("preset,mix", CASES)
def test_stream_logits_and_greedy_tokens(b200, oracle, preset, mix, gen):
    """40-token prompt, then 24 greedy tokens in ONE launch of the kernel."""
    arch, desc, tensors = synth.synth_model(preset, mix, 96)
    gpu = _ctx(b200, desc, tensors, gen)
    assert gpu.path() == PATH[gen]
    ref = oracle.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(40, desc["vocab"])
    want = ref.forward(prompt)
    got = b200.GpuModelWrapper(gpu).forward(prompt, 0)
    assert rel_err(got, want) < TOL
    tok = oracle.argmax_last(want)
    assert tok == oracle.argmax_last(got)
    dev, ms = gpu.decode_greedy(tok, 24)
    seq = []
    for _ in range(24):
        tok = oracle.argmax_last(ref.forward([tok]))
        seq.append(tok)
    assert dev.tolist() == seq and ms > 0
    assert gpu.watchdog() == [0] * 8
    assert gpu.stats()["kernel_launches"] <= 40 + 1 + 1   # one launch per prompt token, one for the 24 greedy tokens
    gpu.close()


def test_stream_is_deterministic_and_matches_first_megakernel(b200):
    """Same tokens and (to rounding: the kernels cut the K sums at different places) the same logits as mega.cuh and
    stream.cuh; two runs of the streamed kernel are bit-identical (tile pieces are added in entry order whichever warp or
    CTA arrives last)."""
    arch, desc, tensors = synth.synth_model("llama-stream-tiny", "Q4_K_M", 128)
    prompt = synth.prompt_tokens(70, desc["vocab"])
    runs = []
    for gen in (2, 2, 0, 1):
        gpu = _ctx(b200, desc, tensors, gen)
        assert gpu.path() == PATH[gen]
        logits = b200.GpuModelWrapper(gpu).forward(prompt, 0)
        toks, _ = gpu.decode_greedy(int(np.argmax(logits)), 16)
        runs.append((np.array(logits), toks.tolist()))
        gpu.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]
    assert rel_err(runs[0][0], runs[2][0]) < 1e-4 and runs[0][1] == runs[2][1]
    assert rel_err(runs[0][0], runs[3][0]) < 1e-4 and runs[0][1] == runs[3][1]


@pytest.mark.parametrize("hidden,ffn,mix", [(2048, 12288, "Q4_K"), (2048, 14336, "Q4_K_M"), (4096, 8192, "Q6_K")])
def test_stream_long_rows_against_first_megakernel(b200, hidden, ffn, mix):
    """Rows of 48 / 56 super-blocks: tiles are cut across CTAs (stream-K ranges, pieces merged through the packet protocol)
    and across the warp pairs of a CTA; random GGUF blocks (no oracle at this size)."""
    from llama_gguf_b200.presets import make_desc
    from llama_gguf_b200.randmodel import random_model

    p = dict(arch="llama", hidden=hidden, n_layers=2, n_heads=hidden // 128, n_kv_heads=hidden // 512, head_dim=128, ffn=ffn, vocab=1024,
             norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False)
    desc, tensors = random_model(p, mix, 256)
    prompt = synth.prompt_tokens(8, desc["vocab"])
    runs = []
    for gen in (2, 0):
        gpu = _ctx(b200, desc, tensors, gen)
        assert gpu.path() == PATH[gen]
        logits = b200.GpuModelWrapper(gpu).forward(prompt, 0)
        toks, _ = gpu.decode_greedy(int(np.argmax(logits)), 12)
        assert gpu.watchdog() == [0] * 8
        runs.append((np.array(logits), toks.tolist()))
        gpu.close()
    assert rel_err(runs[0][0], runs[1][0]) < 1e-4 and runs[0][1] == runs[1][1]


def test_ineligible_shapes_stay_on_the_first_megakernel(b200):
    """Q6_K rows of 4 super-blocks are 840 bytes (not a 16-byte multiple): no tensor map, mega.cuh runs instead."""
    arch, desc, tensors = synth.synth_model("llama-tiny", "Q4_K_M", 64)
    gpu = _ctx(b200, desc, tensors)
    assert gpu.path() == "mega"
    gpu.close()
