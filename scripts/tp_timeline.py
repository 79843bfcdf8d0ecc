"""Per-phase timeline of the streamed megakernel under tensor parallelism (rank 0's CTA 0 stamps after each grid barrier).
usage: torchrun --nproc-per-node N scripts/tp_timeline.py [model] [mix] [prompt_len] [ctx]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import llama_gguf_b200 as B  # noqa: E402
from llama_gguf_b200.parallel import TensorParallelInference  # noqa: E402
from llama_gguf_b200.presets import PRESETS, make_desc  # noqa: E402
from llama_gguf_b200.randmodel import random_model  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
mix = sys.argv[2] if len(sys.argv) > 2 else "Q4_K_M"
plen = int(sys.argv[3]) if len(sys.argv) > 3 else 128
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
preset = PRESETS[model]
desc = make_desc(preset, ctx)
rep = (16 << 20) if model in ("llama-3-70b", "mixtral-8x7b") else None
tp = TensorParallelInference(desc, None, device=lr, feeder=lambda up: random_model(preset, mix, ctx, seed=1236, upload=up, repeat_bytes=rep))
for i in range(plen):
    tp.prefill_token((i * 7919 + 1) % desc["vocab"])
L = B.lib()
h = tp.gpu._h
buf = (C.c_uint64 * 8192)()
L.b200_debug_mega_timeline(h, buf, 8192)  # arm
toks, ms = tp.decode_greedy(1, 16)
n = L.b200_debug_mega_timeline(h, buf, 8192)
if rank == 0:
    t = np.array(buf[:n], dtype=np.float64)
    print(f"{model} {mix} TP={dist.get_world_size()} path {tp.path()}: {ms / 16:.3f} ms/token over 16 tokens at kv_len ~{plen + 16}; {n} stamps")
    t = t[1:]
    d = np.diff(t) / 1000.0
    names = ["QKV gemv", "rope+attn", "O gemv(+reduce)", "gate/up gemv", "down gemv(+reduce)"]
    body = d[:-1]
    body = body[: (len(body) // 5) * 5].reshape(-1, 5)
    for i, nm in enumerate(names):
        print(f"  {nm:20s} mean {body[:, i].mean():7.2f} us  min {body[:, i].min():7.2f}  max {body[:, i].max():7.2f}   x{body.shape[0]} = {body[:, i].sum():8.1f} us")
    print(f"  {'vocab head':20s} {d[-1]:7.2f} us;  sum of phases {d.sum():.1f} us")
tp.close()
dist.barrier()
dist.destroy_process_group()
