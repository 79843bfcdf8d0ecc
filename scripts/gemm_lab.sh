#!/bin/bash
# GPU run of tools/gemm_lab over the shapes that matter (built locally: the binary travels with the snapshot).
mkdir -p gpurun_out
L=gpurun_out/gemm_lab.log
: > $L
run() { echo "== gemm_lab $*" >> $L; timeout 120 tools/gemm_lab "$@" >> $L 2>&1; echo "rc $?" >> $L; }
run 12 14336 4096 32 20 2 0
run 12 14336 4096 32 20 2 16
run 12 14336 4096 32 20 3 0
run 12 14336 4096 32 20 2 1
run 12 14336 4096 32 20 2 15
run 12 14336 4096 2048 10 2 0
run 12 14336 4096 2048 10 2 1
run 12 14336 4096 2048 10 2 8
run 12 1000 4096 300 5
run 12 4096 4096 64 20
run 12 4096 14336 128 20
run 14 4096 14336 32 10
run 8 2048 2048 32 5
cat $L
