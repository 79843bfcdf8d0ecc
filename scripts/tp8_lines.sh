#!/bin/bash
# N-GPU visit: TP / EP bench lines (BASELINE configs[2], [3], [4]); SPECS="model steps mix;..."
set -u
N=${1:-8}
mkdir -p gpurun_out
IFS=";" read -ra SPECA <<< "${SPECS:-llama-3-8b 32;llama-3-70b 16}"
for spec in "${SPECA[@]}"; do
  set -- $spec
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --model $1 --steps $2 --mix ${3:-Q4_K_M} --warmup 4 > gpurun_out/tp${N}_$1.json 2> gpurun_out/tp${N}_$1.err
  echo "bench tp$N $1 exit $?"; grep -a "built and sharded" gpurun_out/tp${N}_$1.err | head -1 | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/tp${N}_$1.json').read().strip().splitlines()[-1])
    print("tp${N} $1: %.1f tok/s  %.3f ms  e2e %.1f  frac %.3f tokens %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("greedy_tokens_head")))
except Exception as e: print("no json", e)
PY
done
