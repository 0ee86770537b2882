#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tokens_tc.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -n 15
for k in tm4 tm3; do
  VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1
  echo "nofold:"; VITCNN_LIB=vit-cnn_b200/csrc/variants/libvitcnn_nofold.so VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1
done
