"""Greedy continuation behind b200_forward (b200_ctx_set_speculation): the device picks argmax after every forward and runs the
next token ahead of the caller.  Whatever the caller then does -- feeds the pick back (hit), feeds another token (miss), calls any
other entry point on the sequence, hits the end of the context -- every result must be bit-identical to the same calls on a
context without it, and within 1e-3 of the oracle."""
import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _pair(b200, preset="llama-stream-tiny", mix="Q4_K_M", ctx=64, max_batch=1):
    arch, desc, tensors = synth.synth_model(preset, mix, ctx, max_batch=max_batch)
    plain = b200.GpuOnlyInference(desc, tensors)
    spec = b200.GpuOnlyInference(desc, tensors)
    spec.set_speculation(True)
    return desc, tensors, plain, spec


def test_greedy_loop_hits_every_step(b200, oracle):
    desc, tensors, plain, spec = _pair(b200)
    assert spec.path() == "stream2" and spec.speculation_stats()["enabled"]
    ref = oracle.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(6, desc["vocab"])
    for t in prompt[:-1]:
        plain.prefill_token(t)
        spec.prefill_token(t)
    ref.forward(prompt[:-1])
    tok = prompt[-1]
    for step in range(24):
        a, b = plain.forward(tok), spec.forward(tok)
        assert np.array_equal(a, b), step
        assert rel_err(b, ref.forward([tok])) < TOL
        assert spec.position() == plain.position() == ref.position()
        tok = oracle.argmax_last(b)
    st = spec.speculation_stats()
    assert st["hits"] == 23 and st["misses"] == 0
    spec.set_speculation(False)                       # drains the token in flight
    assert np.array_equal(plain.forward(tok), spec.forward(tok)) and spec.speculation_stats()["misses"] == 1
    plain.close()
    spec.close()


def test_misses_and_other_entry_points_discard_the_token_in_flight(b200, oracle):
    desc, tensors, plain, spec = _pair(b200, max_batch=2)
    rng = np.random.default_rng(5)
    tok = 7
    for step in range(30):
        kind = step % 6
        if kind == 3:                                  # not the pick: a miss
            tok = int(rng.integers(0, desc["vocab"]))
        if kind == 4:                                  # another entry point between two forwards
            t2 = int(rng.integers(0, desc["vocab"]))
            plain.prefill_token(t2)
            spec.prefill_token(t2)
        if kind == 5:                                  # the other slot, then device-side greedy on this one
            assert np.array_equal(plain.forward(3, 1), spec.forward(3, 1))
            ta, _ = plain.decode_greedy(tok, 2)
            tb, _ = spec.decode_greedy(tok, 2)
            assert ta.tolist() == tb.tolist()
            tok = int(tb[-1])
        a, b = plain.forward(tok), spec.forward(tok)
        assert np.array_equal(a, b), step
        assert plain.position() == spec.position() and plain.position(1) == spec.position(1)
        tok = oracle.argmax_last(b)
    st = spec.speculation_stats()
    assert st["hits"] > 0 and st["misses"] > 0
    spec.reset()
    plain.reset()
    assert spec.position() == 0
    assert np.array_equal(plain.forward(11), spec.forward(11))
    plain.close()
    spec.close()


def test_runs_into_the_end_of_the_context_like_the_plain_path(b200, oracle):
    desc, tensors, plain, spec = _pair(b200, ctx=16)
    tok = 5
    for step in range(16):
        a, b = plain.forward(tok), spec.forward(tok)
        assert np.array_equal(a, b), step
        tok = oracle.argmax_last(b)
    assert spec.position() == 16
    for g in (plain, spec):
        with pytest.raises(b200.InvalidArgument, match="context length"):
            g.forward(tok)
    plain.close()
    spec.close()


def test_paths_without_a_megakernel_ignore_it(b200, oracle):
    desc, tensors, plain, spec = _pair(b200, preset="mixtral-tiny", ctx=32)
    assert spec.path() == "graph" and not spec.speculation_stats()["enabled"]
    tok = 3
    for _ in range(5):
        a, b = plain.forward(tok), spec.forward(tok)
        assert np.array_equal(a, b)
        tok = oracle.argmax_last(b)
    assert spec.speculation_stats() == dict(enabled=False, hits=0, misses=0)
    plain.close()
    spec.close()
