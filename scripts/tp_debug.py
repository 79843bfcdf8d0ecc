"""Two-rank tensor-parallel smoke with progress prints (debugging aid): python scripts/tp_debug.py [world]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(rank, world, port):
    import numpy as np
    import torch
    import torch.distributed as dist

    def say(*a):
        print(f"[rank {rank} +{time.time() - t0:.1f}s]", *a, flush=True)

    t0 = time.time()
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    say("process group up")
    import llama_gguf_b200 as B
    import synth
    from llama_gguf_b200.parallel import TensorParallelInference

    arch, desc, tensors = synth.synth_model("llama-tiny", "Q4_K_M", 64)
    say("model synthesised")
    tp = TensorParallelInference(desc, tensors, device=rank)
    say("TP context finalized")
    prompt = synth.prompt_tokens(6, desc["vocab"])
    for t in prompt[:-1]:
        tp.prefill_token(t)
    say("prefill done")
    logits = tp.forward(prompt[-1])
    say("forward done", logits[:4])
    toks, ms = tp.decode_greedy(int(np.argmax(logits)), 8)
    say("greedy", toks.tolist(), f"{ms / 8:.3f} ms/token")
    if rank == 0:
        import oracle as O

        want = O.OracleModel(desc, tensors).forward(prompt)
        say("rel err vs oracle", synth.rel_err(logits, want))
    tp.close()
    dist.barrier()
    dist.destroy_process_group()
    say("done")


if __name__ == "__main__":
    import torch.multiprocessing as mp

    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    mp.spawn(worker, args=(world, 29731), nprocs=world, join=True)
