#!/bin/bash
# First GPU visit of the second streamed megakernel: small-model parity (oracle + first-generation kernels), an 8B-shape
# 2-layer comparison, then the full bench line and the per-phase timeline.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export B200_LOG=1
run() { echo "=== $*"; timeout 300 "$@" 2>&1 | grep -v "^\[b200\] stream_build: ok" | tail -12; }
run python scripts/s2_debug.py Q4_K_M 5 64
run python scripts/s2_debug.py Q6_K 5 64
run python scripts/s2_debug.py Q8_0 5 64
run python scripts/s2_debug.py Q5_K_M 40 96
run python scripts/s2_debug.py Q4_K_M 40 96 tinyllama-stream-tiny
ORACLE=0 run python scripts/s2_debug.py Q4_K_M 8 512 llama-3-8b 2 4096
ORACLE=0 run python scripts/s2_debug.py Q4_K_M 300 512 llama-3-8b 2 4096
if [ "${1:-}" = "bench" ]; then
  timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err
  echo "bench exit $?"; tail -3 gpurun_out/bench_s2.err; cat gpurun_out/bench_s2.json
  run python scripts/s2_timeline.py llama-3-8b Q4_K_M 128 8192
fi
