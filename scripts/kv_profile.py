"""Decode attention at depth, f32 KV cache vs INT8 KV cache, same per-op decode path for both
(B200_MEGA=0 for the f32 context so that the two attention kernels are the only difference):

    python scripts/kv_profile.py --kv-format int8 --depth 8000 --tokens 3

The KV cache is taken as valid up to --depth as it is (b200_debug_set_position: zero rows -- the kernels' work does not depend on
the values), then --tokens greedy tokens are decoded; prints one JSON line with ms/token (CUDA events).  Run it under
`ncu -k regex:attn --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` for the per-kernel numbers."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kv-format", default="int8", choices=["f32", "int8"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--mix", default="Q4_K_M")
    ap.add_argument("--ctx", type=int, default=8192)
    ap.add_argument("--depth", type=int, default=8000)
    ap.add_argument("--tokens", type=int, default=3)
    args = ap.parse_args()
    os.environ["B200_MEGA"] = "0"
    import llama_gguf_b200 as B
    from llama_gguf_b200.presets import PRESETS, make_desc
    from llama_gguf_b200.randmodel import random_model

    preset = PRESETS[args.model]
    desc = make_desc(preset, args.ctx, 1)
    gpu = B.GpuOnlyInference(desc, None, feeder=lambda up: random_model(preset, args.mix, args.ctx, seed=3, upload=up, repeat_bytes=16 << 20),
                             kv_format=args.kv_format)
    assert gpu.path() == "graph"
    out = {"kv_format": args.kv_format, "model": args.model, "path": gpu.path(), "kv_bytes_per_pos": gpu.stats()["kv_bytes_per_pos"]}
    for depth in (128, args.depth):
        gpu.reset()
        gpu.debug_set_position(depth)
        gpu.decode_greedy(1, 2)                       # warm (graph capture)
        toks, ms = gpu.decode_greedy(1, args.tokens)
        out[f"ms_per_token_at_{depth}"] = ms / args.tokens
    print(json.dumps(out))


if __name__ == "__main__":
    main()
