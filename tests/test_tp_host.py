"""Host-side logic of the tensor-parallel path (CPU only): the sharding plan, shard/merge exactness
(the reference's tests, src/backend/tensor_parallel.rs:330-385) and the handle exchange over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest

import llama_gguf_b200 as B
from llama_gguf_b200 import parallel as TP
from llama_gguf_b200.presets import PRESETS, make_desc, tensor_plan


def test_plan_llama3_8b_divides_for_1_2_4_8():
    d = make_desc(PRESETS["llama-3-8b"], 8192)
    for w in (1, 2, 4, 8):
        p = TP.shard_plan(d, w)
        assert p["n_heads"] * w == 32 and p["n_kv_heads"] * w == 8 and p["ffn"] * w == 14336 and p["vocab"] * w == 128256
        assert p["hidden"] == 4096                      # norms / residual stream replicated
        assert (p["n_heads"] * 128) % 256 == 0          # K of the row-parallel O projection stays block aligned
        assert p["ffn"] % 256 == 0                      # K of the row-parallel down projection


def test_plan_rejects_what_does_not_divide():
    d = make_desc(PRESETS["llama-3-8b"], 8192)
    with pytest.raises(B.InvalidArgument):
        TP.shard_plan(d, 3)
    with pytest.raises(B.InvalidArgument):
        TP.shard_plan(dict(d, n_kv_heads=2), 4)
    with pytest.raises(B.InvalidArgument):
        TP.shard_plan(dict(make_desc(PRESETS["mixtral-8x7b"], 8192), n_experts=6), 4)   # experts must divide
    with pytest.raises(B.InvalidArgument):
        TP.shard_plan(dict(d, vocab=128264), 8)          # vocab / 8 not a multiple of 16


def test_shard_kinds_follow_megatron_split():
    for name, _, _ in tensor_plan(PRESETS["llama-3-8b"], "Q4_K_M"):
        k = TP.shard_kind(name)
        if name.endswith(("attn_q.weight", "attn_k.weight", "attn_v.weight", "ffn_gate.weight", "ffn_up.weight")) or name == "output.weight":
            assert k == 1, name
        elif name.endswith(("attn_output.weight", "ffn_down.weight")):
            assert k == 2, name
        else:
            assert k == 0, name


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("t", [B.Q4_K, B.Q6_K, B.Q8_0, B.F32])
def test_shard_merge_exact(world, t):
    be, bb = B.BLOCK[t]
    k, n = 2048, 64
    rng = np.random.default_rng(world * 100 + t)
    raw = rng.integers(0, 256, size=n * (k // be) * bb, dtype=np.uint8)
    # column-parallel: concatenating the shards in rank order restores the tensor byte for byte
    parts = [TP.shard_tensor("blk.0.ffn_up.weight", t, [k, n], raw, world, r) for r in range(world)]
    assert all(ne == [k, n // world] for ne, _ in parts)
    assert np.array_equal(TP.merge_column([p for _, p in parts]), raw)
    # row-parallel: every row is cut into `world` runs of whole blocks; interleaving them back restores the rows
    parts = [TP.shard_tensor("blk.0.ffn_down.weight", t, [k, n], raw, world, r) for r in range(world)]
    assert all(ne == [k // world, n] for ne, _ in parts)
    rb = (k // be) * bb
    rebuilt = np.concatenate([p.reshape(n, rb // world) for _, p in parts], axis=1).reshape(-1)
    assert np.array_equal(rebuilt, raw)
    # replicated tensors are untouched
    ne, p = TP.shard_tensor("blk.0.attn_norm.weight", B.F32, [k], raw[:k * 4], world, world - 1)
    assert ne == [k] and np.array_equal(p, raw[:k * 4])


def test_row_parallel_partial_sums_add_up():
    """The math the kernels rely on: y = sum_r W[:, K_r] x[K_r] (f32 weights here; the quantised shards split at block
    boundaries, so the same identity holds block for block)."""
    rng = np.random.default_rng(7)
    k, n, world = 1024, 48, 4
    w = rng.standard_normal((n, k)).astype(np.float32)
    x = rng.standard_normal(k).astype(np.float32)
    parts = []
    for r in range(world):
        ne, raw = TP.shard_tensor("blk.0.attn_output.weight", B.F32, [k, n], w, world, r)
        wl = raw.view(np.float32).reshape(n, ne[0])
        parts.append(wl.astype(np.float64) @ x[r * ne[0]:(r + 1) * ne[0]].astype(np.float64))
    assert np.allclose(sum(parts), w.astype(np.float64) @ x.astype(np.float64), rtol=1e-12, atol=1e-9)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        blob = bytes([rank]) * 64                      # stands in for a 64-byte CUDA IPC handle
        got = TP.all_gather_bytes(blob)
        q.put((rank, [g[0] for g in got], [len(g) for g in got]))
    finally:
        dist.destroy_process_group()


def test_handle_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, [0, 1], [64, 64]), (1, [0, 1], [64, 64])]


def test_tp_context_needs_a_gpu_or_fails_loudly():
    """No CPU fallback: without a CUDA device the tensor-parallel constructor raises NotAvailable."""
    if B.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(B.NotAvailable):
        B.GpuOnlyInference(make_desc(PRESETS["llama-3-8b"], 64), {}, parallel=(2, 0), exchange=lambda b: [b, b])


def test_expert_parallel_plan_and_shards():
    """MoE + world > 1 = EXPERT parallel: experts split by the outermost dimension, everything else replicated
    (replaces the single-host rayon loop of src/model/moe.rs:352-361)."""
    d = make_desc(PRESETS["mixtral-8x7b"], 8192)
    plan = TP.shard_plan(d, 8)
    assert plan["expert_parallel"] and plan["n_experts"] == 1 and plan["n_heads"] == d["n_heads"] and plan["vocab"] == d["vocab"]
    rng = np.random.default_rng(3)
    ne = [256, 64, 4]                                  # [K, rows, experts] Q4_K: 144 bytes per row
    data = rng.integers(0, 256, size=4 * 64 * 144, dtype=np.uint8)
    parts = [TP.shard_tensor("blk.0.ffn_gate_exps.weight", B.Q4_K, ne, data, 2, r, expert_parallel=True) for r in range(2)]
    assert parts[0][0] == [256, 64, 2] and parts[1][0] == [256, 64, 2]
    assert np.array_equal(np.concatenate([parts[0][1], parts[1][1]]), data)
    full_ne, full = TP.shard_tensor("blk.0.attn_q.weight", B.Q4_K, [256, 64], data[: 64 * 144], 2, 1, expert_parallel=True)
    assert full_ne == [256, 64] and np.array_equal(full, data[: 64 * 144])   # replicated
