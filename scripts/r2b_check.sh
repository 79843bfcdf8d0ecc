#!/bin/bash
# Round-2 re-entry visit: parity tests, the default bench line, then the N=1 lines of BASELINE configs[3] / configs[4]
# (Llama-3-70B Q4_K_M, Mixtral-8x7B Q5_K_M) at their real shapes.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
if [ "${1:-}" != "nopytest" ]; then
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
fi
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.err | cut -c1-300; cut -c1-1500 gpurun_out/bench.json
for spec in "llama-3-70b Q4_K_M" "mixtral-8x7b Q5_K_M"; do
  set -- $spec
  timeout 1500 python bench.py --model $1 --mix $2 --steps 32 --warmup 4 --depth 0 --batch 0 --prefill-len 0 --no-cpu-baseline \
      > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "bench $1 exit $?"; tail -3 gpurun_out/bench_$1.err | cut -c1-400; cut -c1-1200 gpurun_out/bench_$1.json
done
