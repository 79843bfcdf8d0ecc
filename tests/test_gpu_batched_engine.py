"""Continuous batching: b200_batch_* against a restatement of the reference's BatchedEngine loop
(src/engine_batched.rs:199-411) that steps the ORACLE model one sequence at a time, exactly as the reference does
(step_sequence, :366-411: whole prompt on the first step, then one token per step through Model::forward; greedy pick, last
maximum wins).  The event streams -- token ids, their order across sequences, finish reasons, prompt / completion counts, the
newest-first promotion of pending requests -- must be identical; the GPU side decodes all running sequences in one pass."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


class ReferenceLoop:
    """run_background_loop + create_active_sequence + step_sequence, restated (engine_batched.rs:199-411)."""

    def __init__(self, oracle, desc, tensors, max_batch_size, max_seq_len, max_queue_depth, eos):
        self.O, self.desc, self.tensors = oracle, desc, tensors
        self.max_batch, self.max_seq, self.depth, self.eos = max_batch_size, min(max_seq_len, desc["max_seq_len"]), max_queue_depth, eos
        self.active, self.pending, self.events, self.queue, self.next_id = [], [], [], 0, 1

    def _activate(self, rid, tokens, max_tokens):                        # create_active_sequence (:322-356)
        if not tokens:
            self.events.append(("error", rid, "empty prompt"))
            self.queue -= 1
            return
        toks = list(tokens[: max(0, self.max_seq - 1)])
        self.active.append(dict(id=rid, tokens=toks, prompt_len=len(toks), generated=0, max_tokens=max_tokens,
                                model=self.O.OracleModel(self.desc, self.tensors)))

    def submit(self, tokens, max_tokens):                                # submit (:167-192) + the drain step (:217-239)
        if self.queue >= self.depth:
            raise RuntimeError("queue full")
        self.queue += 1
        rid = self.next_id
        self.next_id += 1
        if len(self.active) < self.max_batch:
            self._activate(rid, tokens, max_tokens)
        else:
            self.pending.append((rid, tokens, max_tokens))
        return rid

    def _step_sequence(self, s):                                         # :366-411
        if s["tokens"] and s["tokens"][-1] == self.eos:
            return None
        if s["generated"] >= s["max_tokens"]:
            return None
        inp = s["tokens"] if s["model"].position() == 0 else s["tokens"][-1:]
        nxt = self.O.argmax_last(s["model"].forward(inp))
        s["tokens"].append(nxt)
        s["generated"] += 1
        return None if nxt == self.eos else nxt

    def step(self):
        out, self.events = self.events, []
        i = 0
        while i < len(self.active):
            s = self.active[i]
            r = self._step_sequence(s)
            if r is None:
                reason = "max_tokens" if s["generated"] >= s["max_tokens"] else "stop"
                self.active.pop(i)
                self.queue -= 1
                out.append(("done", s["id"], reason, s["prompt_len"], s["generated"]))
                continue
            out.append(("token", s["id"], r))
            i += 1
        while len(self.active) < self.max_batch and self.pending:       # pending.pop(): newest first (:291-304)
            self._activate(*self.pending.pop())
        out += self.events
        self.events = []
        return out

    def run(self):
        out = []
        while self.active or self.pending or self.events:
            out += self.step()
        return out


def _requests(desc, n, rng, lens=(1, 24), max_tokens=(1, 9)):
    return [([int(t) for t in rng.integers(0, desc["vocab"], size=int(rng.integers(*lens)))], int(rng.integers(*max_tokens))) for _ in range(n)]


@pytest.mark.parametrize("preset,mix,max_batch,n_req", [("llama-tiny", "Q4_K_M", 3, 8), ("qwen-tiny", "Q4_K_M", 4, 9),
                                                        ("tinyllama-tiny", "Q8_0", 2, 5), ("mixtral-tiny", "Q4_K_M", 3, 6)])
def test_event_stream_equals_the_reference_loop(b200, oracle, preset, mix, max_batch, n_req):
    """Fewer than 8 sequences decode on the exact per-sequence kernels: every token must equal the oracle's."""
    arch, desc, tensors = synth.synth_model(preset, mix, 48, max_batch=max_batch)
    rng = np.random.default_rng(11)
    reqs = _requests(desc, n_req, rng)
    # an EOS that the model actually produces: the most frequent first-generated token of a probe run would do, but any id the
    # greedy decode emits works -- take the first sequence's second generated token so that "stop" and "max_tokens" both occur
    probe = oracle.OracleModel(desc, tensors)
    t = oracle.argmax_last(probe.forward(reqs[0][0]))
    eos = oracle.argmax_last(probe.forward([t]))
    gpu = b200.GpuOnlyInference(desc, tensors)
    eng = b200.BatchedEngine(gpu, max_batch_size=max_batch, max_seq_len=48, max_queue_depth=64, eos_token_id=eos)
    ref = ReferenceLoop(oracle, desc, tensors, max_batch, 48, 64, eos)
    for toks, mt in reqs:
        assert eng.submit(toks, mt) == ref.submit(toks, mt)
    want = ref.run()
    got = eng.run()
    assert got == want
    kinds = {e[0] for e in got} | {e[2] for e in got if e[0] == "done"}
    assert {"token", "done", "max_tokens"} <= kinds
    c = eng.counts()
    assert c["active"] == 0 and c["pending"] == 0 and c["decode_rows"] > 0
    # the slots are free again: a second wave on the same engine gives the same stream as a fresh reference
    ref2 = ReferenceLoop(oracle, desc, tensors, max_batch, 48, 64, eos)
    ids = {}
    for toks, mt in reqs[:3]:
        ids[eng.submit(toks, mt)] = ref2.submit(toks, mt)
    got2 = [(e[0], ids[e[1]]) + tuple(e[2:]) for e in eng.run()]
    assert got2 == ref2.run()
    eng.close()
    gpu.close()


def test_queue_rules(b200, oracle):
    """submit past max_queue_depth -> "queue full" (:173-175); an empty prompt -> Error("empty prompt") (:329-334); a prompt longer
    than max_seq_len - 1 is cut (:336-338); a prompt that ends in EOS finishes with no token (:373-377); max_tokens 0 -> MaxTokens."""
    arch, desc, tensors = synth.synth_model("tinyllama-tiny", "Q8_0", 32, max_batch=2)
    gpu = b200.GpuOnlyInference(desc, tensors)
    with pytest.raises(b200.InvalidArgument, match="max_batch"):
        b200.BatchedEngine(gpu, max_batch_size=3)
    eos = 5
    eng = b200.BatchedEngine(gpu, max_batch_size=2, max_seq_len=16, max_queue_depth=4, eos_token_id=eos)
    ref = ReferenceLoop(oracle, desc, tensors, 2, 16, 4, eos)
    long_prompt = [int(x) for x in (np.arange(40) * 3 + 7) % desc["vocab"] if x != eos]
    reqs = [([], 4), ([9, 8, eos], 4), ([1, 2, 3], 0), (long_prompt, 3), ([4, 4], 2)]
    for toks, mt in reqs:
        assert eng.submit(toks, mt) == ref.submit(toks, mt)
    # the empty prompt left the queue at once; four are queued now (two active, two pending): the queue is full
    assert eng.counts()["active"] == 2 and eng.counts()["pending"] == 2
    with pytest.raises(b200.OperationFailed, match="queue full"):
        eng.submit([1], 1)
    with pytest.raises(RuntimeError, match="queue full"):
        ref.submit([1], 1)
    want, got = ref.run(), eng.run()
    assert got == want
    with pytest.raises(b200.InvalidArgument):
        eng.submit([desc["vocab"]], 1)
    assert ("error", 1, "empty prompt") in got
    assert ("done", 2, "stop", 3, 0) in got and ("done", 3, "max_tokens", 3, 0) in got
    assert [e for e in got if e[0] == "done" and e[1] == 4][0][3] == 15          # cut to max_seq_len - 1
    eng.close()
    gpu.close()


def test_sixteen_sequences_share_the_gemm_pass(b200, oracle):
    """>= 8 running sequences decode through ONE pass of the tcgen05 dequant-GEMMs per step (fp16 operands, logits within 1e-3 of
    the oracle); the launch count shows one pass, not sixteen."""
    nseq = 16
    arch, desc, tensors = synth.synth_model("llama-stream-tiny", "Q4_K_M", 64, max_batch=nseq)
    rng = np.random.default_rng(3)
    reqs = _requests(desc, nseq, rng, lens=(2, 20), max_tokens=(6, 7))
    gpu = b200.GpuOnlyInference(desc, tensors)
    eng = b200.BatchedEngine(gpu, max_batch_size=nseq, max_seq_len=64, max_queue_depth=64, eos_token_id=desc["vocab"] - 1)
    ref = ReferenceLoop(oracle, desc, tensors, nseq, 64, 64, desc["vocab"] - 1)
    for toks, mt in reqs:
        eng.submit(toks, mt)
        ref.submit(toks, mt)
    first = eng.step()                                   # the prompts
    l0 = gpu.stats()["kernel_launches"]
    second = eng.step()                                  # sixteen decode rows in one pass
    assert gpu.stats()["kernel_launches"] - l0 <= 20 * desc["n_layers"] + 8
    got = first + second + eng.run()
    # The GEMM rows are within 1e-3 of the oracle's logits, not bit-equal, so a near-tie may flip a pick and the two streams then
    # part for good.  Tolerance-aware check: replay every sequence's GPU tokens through the oracle -- each pick must be the
    # oracle's argmax or within 2e-3 * max|logit| of it -- and the bookkeeping (counts, reasons, event order per step) must hold.
    by_id = {}
    for e in got:
        by_id.setdefault(e[1], []).append(e)
    exact = 0
    for rid, (toks, mt) in enumerate(reqs, start=1):
        ev = by_id[rid]
        assert [e[0] for e in ev] == ["token"] * mt + ["done"] and ev[-1][2:] == ("max_tokens", len(toks), mt)
        m = oracle.OracleModel(desc, tensors)
        logits = m.forward(toks)
        for e in ev[:-1]:
            assert logits[e[2]] >= logits.max() - 2e-3 * np.abs(logits).max(), (rid, e)
            exact += int(e[2] == oracle.argmax_last(logits))
            logits = m.forward([e[2]])
    assert exact >= 0.9 * sum(mt for _, mt in reqs), "nearly every pick is the oracle's own argmax"
    want = ref.run()
    assert [e[:2] for e in got] == [e[:2] for e in want], "same event kinds for the same requests in the same order"
    # the device pick equals the host pick on the same logits
    for s in range(nseq):
        gpu.reset(s)
        gpu.prefill_token(1 + s, s)
    toks = [3 + s for s in range(nseq)]
    picks = gpu.decode_batch_greedy(list(range(nseq)), toks)
    for s in range(nseq):
        gpu.reset(s)
        gpu.prefill_token(1 + s, s)
    logits = gpu.decode_batch(list(range(nseq)), toks)
    assert picks == [oracle.argmax_last(logits[s]) for s in range(nseq)]
    eng.close()
    gpu.close()
