#!/bin/bash
# quick GPU visit: smoke + short bench with logging (no pytest)
set -u
mkdir -p gpurun_out
export B200_LOG=1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log | cut -c1-400
timeout 600 python bench.py --steps 48 --warmup 8 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "bench exit $?"; grep -a "b200\]" gpurun_out/bench_q.err | head; tail -3 gpurun_out/bench_q.err | cut -c1-600
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
    print("prefill", d.get("prefill")); print("value %.1f tok/s  %.3f ms  e2e %.1f  frac %.3f  tokens %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("greedy_tokens_head")))
except Exception as e: print("no bench json", e)
PY
