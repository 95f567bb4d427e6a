#!/bin/bash
mkdir -p gpurun_out /tmp/gx
F="-gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tests/cuda/gram_tc_check.cu"
i=0
for cfg in "16 4 4 1" "16 4 6 1" "16 4 8 1" "32 2 5 1"; do set -- $cfg; nvcc $F -DGTC_KC=$1 -DGTC_NSTAGE=$2 -DGTC_DEPTH=$3 -DGTC_ISSUER=$4 -o /tmp/gx/c$i 2>/dev/null & nvcc $F -DGTC_KC=$1 -DGTC_NSTAGE=$2 -DGTC_DEPTH=$3 -DGTC_ISSUER=$4 -DGTC_EXP_SKIP=47 -o /tmp/gx/s$i 2>/dev/null & i=$((i+1)); done
wait
i=0
for cfg in "16 4 4 1" "16 4 6 1" "16 4 8 1" "32 2 5 1"; do echo "KC NSTAGE DEPTH ISSUER = $cfg"; /tmp/gx/c$i | tail -1; /tmp/gx/c$i time 2048; echo -n "  skeleton: "; /tmp/gx/s$i time 2048; i=$((i+1)); done 2>&1 | tee gpurun_out/r02o_gram_timing2.txt
