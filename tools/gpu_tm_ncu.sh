#!/bin/bash
# ncu --set full of tokens_tm_kernel (VITCNN_TC_KERNEL=$1), 131 072 patches at P = 11
mkdir -p gpurun_out
k=${1:-tm4}
VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokens_tm_kernel -s 3 -c 1 \
  -o gpurun_out/r2_$k -f python tools/time_tokens.py > gpurun_out/ncu_$k.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_$k.log
