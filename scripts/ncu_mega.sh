#!/bin/bash
# ncu captures of the per-token megakernel (only after the plain run exited 0): launch list + one full capture
set -u
mkdir -p gpurun_out
BARGS="--prompt-len 8 --steps 4 --warmup 3 --ctx 2048 --no-cpu-baseline"
timeout 600 python bench.py $BARGS > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mega_decode -s 9 -c 1 \
    -o gpurun_out/mega_full -f python bench.py $BARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
