"""The single-process group (b200_group_*: one context and one host thread per device INSIDE the library, peers connected with
cudaDeviceEnablePeerAccess) against the oracle and the single-GPU context: what a Rust GpuOnlyInference would call when B200_TP is
set -- no ranks, no other process (TensorParallel trait, src/backend/tensor_parallel.rs:13-32; ShardingPlan :69-106;
MoE models: experts spread over the devices, src/model/moe.rs:321-413 on one host in the reference).
Tolerance: logits within 1e-3 relative of the oracle, greedy tokens identical."""
import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu


def _run(gpu, prompt, n_greedy=8):
    for t in prompt[:-1]:
        gpu.prefill_token(t)
    logits = gpu.forward(prompt[-1])
    toks, _ = gpu.decode_greedy(int(np.argmax(logits)), n_greedy)
    return logits, toks.tolist()


def test_group_of_one_is_the_single_gpu_context(b200, oracle):
    arch, desc, tensors = synth.synth_model("llama-stream-tiny", "Q4_K_M", 64, vocab=1024)
    prompt = synth.prompt_tokens(6, desc["vocab"])
    g = b200.GroupInference(desc, tensors, n_devices=1)
    lg, tg = _run(g, prompt)
    assert g.position() == len(prompt) + 8
    g.reset()
    assert g.position() == 0
    g.close()
    s = b200.GpuOnlyInference(desc, tensors)
    ls, ts = _run(s, prompt)
    s.close()
    assert np.array_equal(lg, ls) and tg == ts
    want = oracle.OracleModel(desc, tensors).forward(prompt)
    assert rel_err(lg, want) < 1e-3


@pytest.mark.parametrize("preset,mix", [("llama-stream-tiny", "Q4_K_M"), ("llama-tiny", "Q8_0"), ("mixtral-tiny", "Q4_K_M")])
def test_group_of_two_devices_in_one_process(b200, oracle, preset, mix):
    if b200.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    arch, desc, tensors = synth.synth_model(preset, mix, 64, vocab=1024 if preset == "llama-stream-tiny" else None)
    prompt = synth.prompt_tokens(6, desc["vocab"])
    g = b200.GroupInference(desc, tensors, n_devices=2)
    assert g.path() == ("graph" if preset == "mixtral-tiny" else "stream2")
    lg, tg = _run(g, prompt)
    g.close()
    s = b200.GpuOnlyInference(desc, tensors)
    ls, ts = _run(s, prompt)
    s.close()
    want = oracle.OracleModel(desc, tensors).forward(prompt)
    assert lg.shape == (desc["vocab"],)
    assert rel_err(lg, want) < 1e-3
    assert rel_err(lg, ls) < 1e-4
    assert tg == ts


def test_group_rejects_bad_device_lists(b200):
    arch, desc, tensors = synth.synth_model("llama-tiny", "Q4_K_M", 32)
    with pytest.raises(b200.InvalidArgument):
        b200.GroupInference(desc, tensors, n_devices=1, devices=[99])
    with pytest.raises((b200.InvalidArgument, b200.NotAvailable)):
        b200.GroupInference(desc, tensors, n_devices=3)
