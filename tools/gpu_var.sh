#!/bin/bash
# same-box A/B of token-kernel variants: tests under the variant library, then timing (default library first)
mkdir -p gpurun_out
for k in ${KERNELS:-tm4}; do
  VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1 | tee -a gpurun_out/tm_time.log
done
for v in "$@"; do
  for k in ${KERNELS:-tm4}; do
    VITCNN_LIB=vit-cnn_b200/csrc/variants/libvitcnn_$v.so VITCNN_TC_KERNEL=$k timeout 300 python -m pytest tests/test_gpu_tokens_tc.py tests/test_gpu_model.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -n 6
    echo "variant $v:" | tee -a gpurun_out/tm_time.log
    VITCNN_LIB=vit-cnn_b200/csrc/variants/libvitcnn_$v.so VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1 | tee -a gpurun_out/tm_time.log
  done
done
