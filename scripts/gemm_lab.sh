#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/gemm_lab.log
: > $L
run() { echo "== gemm_lab $*" >> $L; timeout 120 tools/gemm_lab "$@" >> $L 2>&1; echo "rc $?" >> $L; }
run 12 14336 4096 2048 10 2 0
run 12 14336 4096 2048 10 3 0
run 12 4096 14336 2048 10 3 0
run 14 4096 14336 2048 10 2 0
run 14 4096 14336 2048 10 3 0
run 12 4096 4096 512 10 2 0
run 12 4096 4096 512 10 3 0
grep "gemm_lab\|umma2\|check" $L
