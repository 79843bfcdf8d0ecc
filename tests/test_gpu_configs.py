"""Parity at the shapes BASELINE.json names, not miniatures (VERDICT r1 "parity at real config shapes"), plus the tie rules of
the greedy pick (src/main.rs:1816-1821: LAST maximal index) and of the MoE router (src/model/moe.rs:168: stable sort, lowest
index first) exercised on the device, the Backend methods exported this round and the opt-in prefill queue.
Tolerance: logits within 1e-3 relative (max|a-b| / max|b|), greedy token sequences identical (north star)."""
import os

import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def test_config0_qwen25_05b_q4km_128_prompt_64_greedy(b200, oracle):
    """BASELINE configs[0] verbatim: Qwen2.5-0.5B architecture (H = 896: Q4_K -> Q5_0 and Q6_K -> Q8_0 fallback types, NeoX RoPE,
    q/k/v biases, tied 151 936-row head), Q4_K_M mix, 128-token prompt + 64 greedy tokens, batch 1."""
    arch, desc, tensors = synth.synth_model("qwen2.5-0.5b", "Q4_K_M", 256)
    gpu = b200.GpuOnlyInference(desc, tensors)
    ref = oracle.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(128, desc["vocab"])
    for t in prompt[:-1]:
        gpu.prefill_token(t)          # exact token-by-token path (what GpuModelWrapper::forward does, backend/mod.rs:343-346)
    got = gpu.forward(prompt[-1])
    want = ref.forward(prompt)
    assert rel_err(got, want) < TOL
    tok = oracle.argmax_last(want)
    assert tok == oracle.argmax_last(got)
    dev, _ = gpu.decode_greedy(tok, 64)
    seq = []
    for _ in range(64):
        tok = oracle.argmax_last(ref.forward([tok]))
        seq.append(tok)
    assert dev.tolist() == seq
    assert gpu.position() == 128 + 64 == ref.position()
    gpu.close()


def test_llama3_8b_shape_two_layers_full_vocab_head(b200, oracle):
    """Llama-3-8B shapes (H = 4096, 32 / 8 heads of 128, I = 14336, the full 128 256-row Q6_K head), 2 layers, Q4_K_M mix: the
    streamed megakernel at the bench's real tile counts (stream-K ranges, tiles cut across CTAs) against the oracle."""
    arch, desc, tensors = synth.synth_model("llama-3-8b", "Q4_K_M", 64, n_layers=2)
    gpu = b200.GpuOnlyInference(desc, tensors)
    assert gpu.path() == "stream2"
    ref = oracle.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(6, desc["vocab"])
    want = ref.forward(prompt)
    got = b200.GpuModelWrapper(gpu).forward(prompt, 0)
    assert rel_err(got, want) < TOL
    tok = oracle.argmax_last(want)
    assert tok == oracle.argmax_last(got)
    dev, _ = gpu.decode_greedy(tok, 6)
    seq = []
    for _ in range(6):
        tok = oracle.argmax_last(ref.forward([tok]))
        seq.append(tok)
    assert dev.tolist() == seq
    assert gpu.watchdog() == [0] * 8
    gpu.close()


@pytest.mark.parametrize("gen", [2, 1, 0, -1])
def test_greedy_tie_last_index_wins(b200, oracle, gen):
    """Every row of the vocab head identical -> every logit equal -> the LAST index must win (src/main.rs:1816-1821), on every
    decode path: stream2 (candidates collected in the head's epilogue, merged across warps and CTAs), stream, mega, graph."""
    arch, desc, tensors = synth.synth_model("llama-stream-tiny", "Q4_K_M", 32)
    t, ne, data = tensors["output.weight"]
    rows = np.asarray(data).reshape(desc["vocab"], -1)
    rows[:] = rows[7]
    tensors["output.weight"] = (t, ne, rows.reshape(-1))
    env = {2: {}, 1: {"B200_STREAM2": "0"}, 0: {"B200_STREAM2": "0", "B200_STREAM": "0"}, -1: {"B200_MEGA": "0"}}[gen]
    os.environ.update(env)
    try:
        gpu = b200.GpuOnlyInference(desc, tensors)
    finally:
        for k in env:
            os.environ.pop(k, None)
    assert gpu.path() == {2: "stream2", 1: "stream", 0: "mega", -1: "graph"}[gen]
    ref = oracle.OracleModel(desc, tensors)
    logits = gpu.forward(3)
    assert oracle.argmax_last(ref.forward([3])) == desc["vocab"] - 1          # the reference: exact ties, last index
    if gen == 2:
        # stream-K cuts the tiles of the head at CTA boundaries, so identical rows in DIFFERENT tiles are summed in different
        # groupings (4e-7 apart); rows of one tile share every cut: the maximum is attained by a whole tile's worth of rows
        assert rel_err(logits, np.full_like(logits, logits[0])) < 1e-5
        assert int(np.sum(logits == logits.max())) >= 8
    else:
        assert np.all(logits == logits[0])
    want = oracle.argmax_last(logits)                                         # LAST of the equal maxima of the device's own logits
    gpu.reset()
    dev, _ = gpu.decode_greedy(3, 1)                                          # same token, same position: the same logits, picked on the device
    assert dev.tolist() == [want]
    if gen != 2:
        assert want == desc["vocab"] - 1
    gpu.close()


def test_router_tie_lowest_expert_first(b200, oracle):
    """All router rows identical -> every expert logit equal -> the stable descending sort keeps experts 0 and 1 (moe.rs:168), weights 1/2 each."""
    arch, desc, tensors = synth.synth_model("mixtral-tiny", "Q4_K_M", 32)
    for name in list(tensors):
        if name.endswith("ffn_gate_inp.weight"):
            t, ne, data = tensors[name]
            w = np.asarray(data, dtype=np.float32).reshape(desc["n_experts"], -1).copy()
            w[:] = w[2]
            tensors[name] = (t, ne, w.reshape(-1))
    gpu = b200.GpuOnlyInference(desc, tensors)
    ref = oracle.OracleModel(desc, tensors)
    for tok in synth.prompt_tokens(5, desc["vocab"]):
        assert rel_err(gpu.forward(tok), ref.forward([tok])) < TOL
    gpu.close()


def test_backend_matmul_matvec_matvec_q(b200, oracle):
    """Backend::matmul / matvec / matvec_q (src/backend/mod.rs:90, 93, 107; cpu/ops.rs:429-487, 531-575, 922-946)."""
    be = b200.CudaB200Backend()
    rng = np.random.default_rng(5)
    a = rng.standard_normal((7, 33)).astype(np.float32)
    b = rng.standard_normal((33, 12)).astype(np.float32)
    want = np.zeros((7, 12), dtype=np.float32)
    for kk in range(33):                       # matmul_simple: k summed in order, no FMA
        want = (want + a[:, kk:kk + 1] * b[kk:kk + 1, :]).astype(np.float32)
    assert np.array_equal(be.matmul(a, b), want)
    x = rng.standard_normal(33).astype(np.float32)
    assert rel_err(be.matvec(a, x), a.astype(np.float64) @ x.astype(np.float64)) < 1e-6
    k, m = 512, 19
    w = (0.05 * rng.standard_normal(k * m)).astype(np.float32)
    raw = oracle.quantize(12, w)
    xv = rng.standard_normal(k).astype(np.float32)
    got = be.matvec_q(raw, 12, xv, m, k)
    assert rel_err(got, oracle.vec_mat_q(12, raw, xv, m)) < 1e-5
    with pytest.raises(b200.ShapeMismatch):
        be.matmul(a, b[:-1])
    with pytest.raises(b200.InvalidArgument):
        be.matvec(a[0], x)


def test_prefill_queue_reaches_the_gemm_prefill_through_prefill_token(b200, oracle):
    """B200_PREFILL_QUEUE=1: prefill_token only queues (position() counts the queue); the next forward runs the queue as ONE
    tcgen05 GEMM prefill (fp16 operands: 3e-3) -- the path an unchanged GpuModelWrapper::forward takes (backend/mod.rs:343-348)."""
    arch, desc, tensors = synth.synth_model("llama-stream-tiny", "Q4_K_M", 128)
    os.environ["B200_PREFILL_QUEUE"] = "1"
    try:
        gpu = b200.GpuOnlyInference(desc, tensors)
    finally:
        os.environ.pop("B200_PREFILL_QUEUE", None)
    ref = oracle.OracleModel(desc, tensors)
    prompt = synth.prompt_tokens(70, desc["vocab"])
    l0 = gpu.stats()["kernel_launches"]
    for t in prompt[:-1]:
        gpu.prefill_token(t)
    assert gpu.position() == 69 and gpu.stats()["kernel_launches"] == l0     # nothing has run yet
    got = gpu.forward(prompt[-1])
    want = ref.forward(prompt)
    assert gpu.position() == 70
    assert rel_err(got, want) < 3e-3
    assert oracle.argmax_last(got) == oracle.argmax_last(want)
    gpu.reset()
    assert gpu.position() == 0
    gpu.prefill_token(1)
    gpu.reset()                                                              # a reset drops the queue
    assert gpu.position() == 0
    gpu.close()


def test_llama3_8b_shape_int8_kv_from_a_gguf_file_and_the_batched_engine(b200, oracle, tmp_path):
    """The rows added last at Llama-3-8B shapes, 2 layers, full 128 256-row head: the model is written to a
    GGUF file, loaded by the library's own reader (b200_ctx_create_from_gguf) with the INT8 KV cache, decoded against the oracle's
    int8-cache model; then an f32 context of the same file runs four requests through the continuous-batching engine against the
    reference loop's restatement (exact path: < 8 rows)."""
    from llama_gguf_b200 import gguf_io
    from test_gpu_batched_engine import ReferenceLoop

    arch, desc, tensors = synth.synth_model("llama-3-8b", "Q4_K_M", 64, n_layers=2, max_batch=4)
    path = os.path.join(tmp_path, "llama3-8b-2l.gguf")
    gguf_io.write_gguf(path, arch, desc, tensors)
    gpu = b200.GpuOnlyInference.from_gguf(path, max_seq_len=64, max_batch=4, kv_format="int8")
    assert gpu.kv_format() == "int8" and gpu.load_stats["tensors_loaded"] == len(tensors)
    ref = oracle.OracleModel(desc, tensors, kv_format="int8")
    prompt = synth.prompt_tokens(20, desc["vocab"])
    want = ref.forward(prompt)
    assert rel_err(gpu.prefill(prompt), want) < TOL
    tok = oracle.argmax_last(want)
    for _ in range(6):
        want = ref.forward([tok])
        got = gpu.forward(tok)
        assert rel_err(got, want) < TOL and oracle.argmax_last(got) == oracle.argmax_last(want)
        tok = oracle.argmax_last(want)
    gpu.close()
    gpu = b200.GpuOnlyInference.from_gguf(path, max_seq_len=64, max_batch=4)
    assert gpu.path() == "stream2"
    eng = b200.BatchedEngine(gpu, max_batch_size=3, max_seq_len=64, max_queue_depth=8, eos_token_id=desc["vocab"] - 1)
    loop = ReferenceLoop(oracle, desc, tensors, 3, 64, 8, desc["vocab"] - 1)
    rng = np.random.default_rng(2)
    for _ in range(4):
        toks = [int(t) for t in rng.integers(0, desc["vocab"], size=int(rng.integers(2, 9)))]
        mt = int(rng.integers(2, 5))
        assert eng.submit(toks, mt) == loop.submit(toks, mt)
    assert eng.run() == loop.run()
    eng.close()
    gpu.close()
