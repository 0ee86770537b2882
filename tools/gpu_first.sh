#!/bin/bash
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
echo "=== gather"; timeout 600 python -m pytest tests/test_gpu_gather.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15
echo "=== conv simt + pack"; timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --tb=short -p no:cacheprovider -k "simt or pack" 2>&1 | tail -15
for c in 0 1 4; do for f in 0 1; do echo "=== probe conv case $c flags $f"; timeout 120 python tools/probe_conv.py $c $f 2>&1 | tail -22; done; done
echo "=== conv tcgen05"; timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --tb=line -p no:cacheprovider -k "tcgen05" 2>&1 | tail -15
echo "=== model"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25
echo "=== scene"; timeout 900 python -m pytest tests/test_gpu_scene.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
