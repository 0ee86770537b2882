#!/bin/bash
# training-step evidence: tests, bench (training leg), ncu launch list and full captures
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 500 python -m pytest tests/test_gpu_train.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5
PROF="python bench.py --no-infer --steps 1 --warmup 3"
$PROF > gpurun_out/train_plain.log 2>&1; tail -1 gpurun_out/train_plain.log
$PROF > gpurun_out/train_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/train_launches.csv $PROF > gpurun_out/ncu_train_list.log 2>&1
echo "ncu list rc=$?"
$PROF > gpurun_out/train_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"wgrad_sps_tc|transformer_bwd" -s 17 -c 17 -o gpurun_out/prof_train $PROF > gpurun_out/ncu_train_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_train_full.log
