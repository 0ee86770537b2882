#!/bin/bash
# tokens_tm_kernel: parity tests with three / four patches in flight, then same-box timing against tokens_tc_kernel
mkdir -p gpurun_out
for k in tm3 tm4; do
  VITCNN_TC_KERNEL=$k timeout 300 python -m pytest tests/test_gpu_tokens_tc.py tests/test_gpu_model.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tm_test_$k.log 2>&1
  echo "== tokens tests $k rc=$?"; tail -n 4 gpurun_out/tm_test_$k.log
done
for k in ${KERNELS:-tc tm3 tm4}; do
  VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1 | tee -a gpurun_out/tm_time.log
done
for v in "$@"; do
  for k in tm3 tm4; do
    echo "variant $v:" | tee -a gpurun_out/tm_time.log
    VITCNN_LIB=vit-cnn_b200/csrc/variants/libvitcnn_$v.so VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1 | tee -a gpurun_out/tm_time.log
  done
done
