#!/bin/bash
# round-2 ncu evidence for the prefill / batched-decode paths: launch lists (gpu__time_duration) + one full capture of the
# persistent dequant-GEMM (gate shape, T = 2048) and of the tensor-core attention
set -u
mkdir -p gpurun_out
timeout 300 python scripts/prefill_profile.py 2048 || { echo "plain run failed"; exit 1; }
timeout 300 python scripts/batch_profile.py || { echo "plain batch run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_prefill_launches.csv python scripts/prefill_profile.py 2048 > /dev/null 2>&1
echo "prefill launch list exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_batch_launches.csv python scripts/batch_profile.py > /dev/null 2>&1
echo "batch launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dequant_gemm_umma2 -s 40 -c 2 -o gpurun_out/r02_gemm2_full -f python scripts/prefill_profile.py 2048 > gpurun_out/ncu_gemm2.log 2>&1
echo "gemm2 full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_umma -s 2 -c 1 -o gpurun_out/r02_attn_full -f python scripts/prefill_profile.py 2048 > gpurun_out/ncu_attn.log 2>&1
echo "attn full exit $?"
ls -la gpurun_out/*.ncu-rep
