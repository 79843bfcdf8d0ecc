"""GGUF -> HBM load time: a random-init GGUF of a named architecture is written to local disk, then loaded
into a finalized context by every path the package has, wall clock from the path to a context that can decode:

  native            b200_ctx_create_from_gguf: mmap + parse + pinned double-buffered upload inside the library (csrc/gguf_load.cuh)
  native_1thread    the same with one host thread filling the staging buffers
  native_8threads   ... with eight
  native_unstaged   cudaMemcpy straight from the mapping (pageable memory: the driver stages it itself)
  python_shim       round 1's path: gguf-py reader -> numpy arrays -> b200_ctx_upload_tensor per tensor (gguf_io.load_gguf)

Every context decodes the same 4-token prompt and the logits must agree bit for bit.  One JSON line on stdout.
Usage (GPU box): python scripts/load_bench.py [--model llama-3-8b] [--mix Q4_K_M] [--dir /tmp]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--mix", default="Q4_K_M")
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--skip-python", action="store_true")
    args = ap.parse_args()

    import llama_gguf_b200 as B
    from llama_gguf_b200 import gguf_io
    from llama_gguf_b200.presets import PRESETS
    from llama_gguf_b200.randmodel import random_model

    preset = PRESETS[args.model]
    path = os.path.join(args.dir, f"{args.model}-{args.mix}.gguf")
    t0 = time.perf_counter()
    desc, tensors = random_model(preset, args.mix, args.ctx, seed=7, repeat_bytes=16 << 20)
    gguf_io.write_gguf(path, preset["arch"], desc, tensors)
    del tensors
    size = os.path.getsize(path)
    print(f"wrote {path}: {size / 1e9:.2f} GB in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    prompt = [(i * 7919 + 1) % desc["vocab"] for i in range(4)]

    def run(label, env, fn):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            t = time.perf_counter()
            gpu = fn()
            wall = time.perf_counter() - t
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        logits = gpu.prefill(prompt)
        st = getattr(gpu, "load_stats", None)
        gpu.close()
        rec = {"wall_s": wall, "GBps_wall": size / wall / 1e9}
        if st:
            rec.update(load_loop_s=st["seconds"], GBps_load_loop=st["GBps"], device_bytes=st["device_bytes"],
                       tensors_loaded=st["tensors_loaded"])
        print(f"{label}: {rec}", file=sys.stderr)
        return rec, logits

    native = lambda: B.GpuOnlyInference.from_gguf(path, max_seq_len=args.ctx)   # noqa: E731
    out = {"model": args.model, "mix": args.mix, "file_bytes": size, "host_cores": os.cpu_count(), "runs": {}}
    run("warm (page cache, CUDA context, first pinned allocation)", {}, native)
    ref = None
    for label, env in (("native", {}), ("native_1thread", {"B200_LOAD_THREADS": "1"}), ("native_8threads", {"B200_LOAD_THREADS": "8"}),
                       ("native_unstaged", {"B200_LOAD_STAGED": "0"})):
        rec, logits = run(label, env, native)
        out["runs"][label] = rec
        if ref is None:
            ref = logits
        assert np.array_equal(ref, logits), label
    if not args.skip_python:
        def shim():
            arch, d2, t2 = gguf_io.load_gguf(path)
            return B.GpuOnlyInference.from_model((d2, t2), args.ctx)
        rec, logits = run("python_shim", {}, shim)
        out["runs"]["python_shim"] = rec
        assert np.array_equal(ref, logits)
    out["logits_identical"] = True
    os.remove(path)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
