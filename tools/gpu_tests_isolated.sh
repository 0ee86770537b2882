#!/bin/bash
# Run every GPU test file in its own process with a timeout, so a trapped kernel (sticky CUDA
# error) in one file cannot take the others down.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
rc_all=0
for f in tests/test_gpu_gather.py tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_scene.py "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $f rc=$rc"; tail -n 25 gpurun_out/$name.log
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
