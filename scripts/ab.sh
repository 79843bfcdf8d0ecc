#!/bin/bash
# A/B sweep of engine knobs (environment variables) with the short bench: scripts/ab.sh "VAR=a VAR=b ..." (one run per token)
set -u
mkdir -p gpurun_out
BARGS="--steps 48 --warmup 8 --no-cpu-baseline"
for cfg in "$@"; do
  out=$(env $cfg timeout 600 python bench.py $BARGS 2> gpurun_out/ab.err | tail -1)
  echo "$cfg -> $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f tok/s  %.3f ms  e2e %.1f  tok %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["greedy_tokens_head"][:3]))' 2>&1 | tail -1)"
done
