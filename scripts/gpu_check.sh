#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, then (only if the plain bench exited 0)
# the ncu launch list and one full capture of the GEMV kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
BARGS="--prompt-len 8 --steps 4 --warmup 3 --no-cpu-baseline"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${1:-}" = "ncu" ]; then
  timeout 600 python bench.py $BARGS > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2330 -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 600 python bench.py $BARGS > gpurun_out/bench_small2.json 2> gpurun_out/bench_small2.err &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 1600 -c 6 \
      -o gpurun_out/gemv_full -f python bench.py $BARGS > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
  ls -la gpurun_out/
fi
