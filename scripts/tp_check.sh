#!/bin/bash
# N-GPU visit: TP parity tests, then the TP bench line with the REDUCE phase folded (default) and as a phase of its own
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tp${N}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/tp${N}_pytest.log
for fold in ${FOLDS:-0}; do
  B200_TP_LL=${LL:-1} B200_TP_FOLD=$fold timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 32 --warmup 4 ${BENCH_ARGS:-} > gpurun_out/tp${N}_fold${fold}.json 2> gpurun_out/tp${N}_fold${fold}.err
  echo "bench tp$N fold=$fold exit $?"; tail -2 gpurun_out/tp${N}_fold${fold}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/tp${N}_fold${fold}.json').read().strip().splitlines()[-1])
    print("tp${N} fold=${fold}: %.1f tok/s  %.3f ms  tokens %s" % (d["value"], d["ms_per_step"], d.get("greedy_tokens_head")))
except Exception as e: print("no json", e)
PY
done
