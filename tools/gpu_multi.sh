#!/bin/bash
# bench.py on N GPUs of one box, launched the way the driver does (usage: gpurun --gpus N -- bash tools/gpu_multi.sh N)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "rc=$?"; tail -c 600 gpurun_out/r02_bench_n$N.json | head -c 300; echo
