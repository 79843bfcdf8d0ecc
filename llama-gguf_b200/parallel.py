"""Tensor-parallel (dense models) and expert-parallel (MoE models) host side of the cuda-b200 backend: one process per GPU
(torch.distributed for the plumbing), Megatron-style column / row shards of the quantised weights with the all-reduce inside
the per-token kernel, or experts spread over the GPUs (everything else replicated) with the combine through (value, epoch)
packets -- both over NVLink peer memory, no NCCL call on the data path.

Mirrors the reference's ShardingPlan (src/backend/tensor_parallel.rs:69-106) and replaces its gRPC
all-reduce-via-rank-0 (src/distributed/tensor_parallel_distributed.rs:135-187).  `shard_plan` / `shard_tensor` /
`merge_column` are the pure host logic (CPU-testable, tests/test_tp_host.py); the C library applies the same split
when a rank uploads a full tensor (csrc/engine.cu: tp_shard_kind, b200_ctx_upload_tensor)."""
import numpy as np

from . import BLOCK, GpuOnlyInference, InvalidArgument

COLUMN_SUFFIXES = ("attn_q.weight", "attn_k.weight", "attn_v.weight", "attn_q.bias", "attn_k.bias", "attn_v.bias",
                   "ffn_gate.weight", "ffn_up.weight")
ROW_SUFFIXES = ("attn_output.weight", "ffn_down.weight")


def shard_kind(name: str) -> int:
    """0 = replicated, 1 = column-parallel (split output rows), 2 = row-parallel (split K)."""
    if name == "output.weight" or name.endswith(COLUMN_SUFFIXES):
        return 1
    if name.endswith(ROW_SUFFIXES):
        return 2
    return 0


def shard_plan(desc: dict, world: int) -> dict:
    """Local dimensions of one rank, or InvalidArgument if the model does not divide (tensor_parallel.rs:69-106)."""
    if world not in (1, 2, 4, 8):
        raise InvalidArgument("tensor-parallel world size must be 1, 2, 4 or 8")
    moe = desc.get("n_experts", 0) > 0
    for k in (() if moe and world > 1 else ("n_heads", "n_kv_heads", "ffn", "vocab")):
        if desc[k] % world:
            raise InvalidArgument(f"{k}={desc[k]} is not divisible by the world size {world}")
    if world > 1 and not moe and (desc["vocab"] // world) % 16:
        raise InvalidArgument("vocab / world_size must be a multiple of 16")
    if world > 1 and desc.get("n_experts", 0) > 0:
        # MoE: EXPERT parallel (attention / norms / router / head replicated, expert e on rank e // (E / world))
        if desc["n_experts"] % world:
            raise InvalidArgument(f"n_experts={desc['n_experts']} is not divisible by the world size {world}")
        return {"n_heads": desc["n_heads"], "n_kv_heads": desc["n_kv_heads"], "ffn": desc["ffn"], "vocab": desc["vocab"],
                "hidden": desc["hidden"], "n_experts": desc["n_experts"] // world, "expert_parallel": True}
    return {"n_heads": desc["n_heads"] // world, "n_kv_heads": desc["n_kv_heads"] // world, "ffn": desc["ffn"] // world,
            "vocab": desc["vocab"] // world, "hidden": desc["hidden"]}


def shard_tensor(name: str, ggml_type: int, ne, data: np.ndarray, world: int, rank: int, expert_parallel: bool = False):
    """(ne_local, bytes_local) of `rank`'s shard of a tensor in GGUF block layout (rows of ne[0]/bs blocks).
    expert_parallel: only the *_exps tensors are split (by expert: the outermost dimension), everything else is replicated."""
    raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    if expert_parallel and world > 1:
        ne = list(ne)
        if "_exps.weight" not in name:
            return ne, raw
        if len(ne) != 3 or ne[2] % world:
            raise InvalidArgument(f"{name}: experts not divisible by the world size")
        n = raw.size // world
        ne[2] //= world
        return ne, raw[rank * n:(rank + 1) * n]
    kind = shard_kind(name) if world > 1 else 0
    be, bb = BLOCK[ggml_type]
    ne = list(ne)
    if kind == 0:
        return ne, raw
    if kind == 1:
        dim = 0 if len(ne) == 1 else 1
        if ne[dim] % world:
            raise InvalidArgument(f"{name}: rows not divisible by the world size")
        n = raw.size // world
        ne[dim] //= world
        return ne, raw[rank * n:(rank + 1) * n]
    nb = ne[0] // be
    if nb % world:
        raise InvalidArgument(f"{name}: K blocks not divisible by the world size")
    rows = raw.size // (nb * bb)
    sl = nb // world * bb
    out = raw.reshape(rows, nb * bb)[:, rank * sl:(rank + 1) * sl]
    ne[0] //= world
    return ne, np.ascontiguousarray(out).reshape(-1)


def merge_column(parts):
    """Inverse of a column-parallel split (rows concatenated in rank order)."""
    return np.concatenate([np.asarray(p).reshape(-1) for p in parts])


def all_gather_bytes(blob: bytes, group=None):
    """The `exchange` callable for GpuOnlyInference: all-gather of equal-sized byte strings over torch.distributed
    (works on gloo and nccl: the payload is a CPU uint8 tensor on gloo, a CUDA one on nccl)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


class TensorParallelInference:
    """GpuInference over `world` ranks.  Every rank constructs it with the same full model; forward() returns the full
    logits on every rank (slices gathered over torch.distributed), decode_greedy() never leaves the device: the
    per-rank argmax candidates are exchanged through peer memory."""

    def __init__(self, desc: dict, tensors, device=None, feeder=None, group=None):
        import torch
        import torch.distributed as dist

        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        shard_plan(desc, self.world)
        if device is None:
            device = torch.cuda.current_device()
        self.gpu = GpuOnlyInference(desc, tensors, device=device, feeder=feeder, parallel=(self.world, self.rank),
                                    exchange=lambda b: all_gather_bytes(b, group))
        self.vocab = desc["vocab"]
        self.expert_parallel = self.gpu.expert_parallel   # MoE: every rank holds the full logits (replicated head), nothing to gather
        dist.barrier(group)

    def _gather(self, local):
        import torch
        import torch.distributed as dist

        if self.world == 1 or self.expert_parallel:
            return local
        backend = dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.from_numpy(local).to(dev)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t, group=self.group)
        return merge_column([o.cpu().numpy() for o in outs])

    def forward(self, token_id):
        return self._gather(self.gpu.forward(token_id))

    def prefill_token(self, token_id):
        self.gpu.prefill_token(token_id)

    def decode_greedy(self, first_token, n_steps):
        return self.gpu.decode_greedy(first_token, n_steps)

    def reset(self):
        self.gpu.reset()

    def position(self):
        return self.gpu.position()

    def stats(self):
        return self.gpu.stats()

    def path(self):
        return self.gpu.path()

    def close(self):
        self.gpu.close()
