#!/bin/bash
# GPU run of tools/gemm_lab over the shapes that matter (built locally: the binary travels with the snapshot).
mkdir -p gpurun_out
L=gpurun_out/gemm_lab.log
: > $L
run() { echo "== gemm_lab $*" >> $L; timeout 120 tools/gemm_lab "$@" >> $L 2>&1; echo "rc $?" >> $L; }
run 14 4096 14336 2048 5
run 14 1024 4096 2048 5
run 14 4096 14336 32 10
run 14 1000 2048 300 5
run 12 14336 4096 2048 10
run 12 4096 14336 32 20
grep -v "role\|^rc\|plan" $L
