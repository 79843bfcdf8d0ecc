#!/bin/bash
# ncu full capture (with source) of ONE launch of the streamed megakernel: 2 greedy tokens of Llama-3-8B Q4_K_M
set -u
mkdir -p gpurun_out
BARGS="--prompt-len 8 --steps 2 --warmup 3 --ctx 2048 --no-cpu-baseline --prefill-len 0"
timeout 600 python bench.py $BARGS > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || { echo "plain run failed"; tail -3 gpurun_out/bench_small.err; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"stream2_decode|stream_decode|mega_decode" -s 9 -c 1 \
    -o gpurun_out/stream_full -f python bench.py $BARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
