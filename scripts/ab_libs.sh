#!/bin/bash
# A/B of library variants (build/variants/*.so, selected with B200_LIB): the quick 8B bench line of each, twice, interleaved
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = "main" ]; then unset B200_LIB; else export B200_LIB=$PWD/build/variants/$lib.so; fi
  timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --depth 0 --batch 0 --prefill-len 0 > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err
  python - "$lib" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/ab_{lib}.json').read().strip().splitlines()[-1])
    print("%-10s value %.1f tok/s  %.4f ms  e2e %.1f" % (lib, d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(lib, "failed", e)
PY
done
done
