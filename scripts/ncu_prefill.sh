#!/bin/bash
# ncu evidence for the GEMM prefill: launch list of one 2048-token prefill (2-layer model, Llama-3-8B shapes) and a full
# capture of the gate / up / down dequant-GEMMs of its first layer
set -u
mkdir -p gpurun_out
timeout 300 python scripts/prefill_profile.py 2048 || { echo "plain run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prefill_launches.csv python scripts/prefill_profile.py 2048 > /dev/null 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dequant_gemm_umma -s 18 -c 3 -o gpurun_out/prefill_gemm_full -f python scripts/prefill_profile.py 2048 > gpurun_out/ncu_prefill.log 2>&1
echo "full exit $?"; ls -la gpurun_out/prefill_gemm_full.ncu-rep
