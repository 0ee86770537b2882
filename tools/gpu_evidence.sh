#!/bin/bash
# Round evidence: all gpu tests, smoke, bench (both arms), ncu launch list and full captures of every kernel of
# one inference step (B200_PROFILING.md recipe).  TRAIN=1 adds the training-step launch list and captures.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
PROF="python bench.py --steps 1 --warmup 3 --windows 65536 --no-cpu --no-e2e --no-train"
$PROF > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
# full captures (no source import: gpurun brings back at most 64 MiB) of one launch of every kernel of the second
# step of the shared-stem scene path.  Per step: 9 conv-1 variants (pair kernel), 25 conv-2 variants + 2 per-window
# conv 3 (single-CTA kernel), then per chunk LiDAR strip gather, variant gather, LiDAR stem, token stage + cls tail.
NCU="ncu --set full --clock-control none"
$PROF > gpurun_out/prof_plain2.log 2>&1 || echo "plain run failed"
$NCU -k regex:conv_sps_tc2_kernel -s 13 -c 1 -o gpurun_out/prof_conv1_interior -f $PROF > gpurun_out/ncu_full_a.log 2>&1; echo "ncu conv1 rc=$?"
$NCU -k regex:conv_sps_tc_kernel -s 27 -c 1 -o gpurun_out/prof_conv2_corner -f $PROF > gpurun_out/ncu_full_b.log 2>&1; echo "ncu conv2 corner rc=$?"
$NCU -k regex:conv_sps_tc_kernel -s 39 -c 1 -o gpurun_out/prof_conv2_interior -f $PROF > gpurun_out/ncu_full_c.log 2>&1; echo "ncu conv2 interior rc=$?"
$NCU -k regex:conv_sps_tc_kernel -s 52 -c 1 -o gpurun_out/prof_conv3 -f $PROF > gpurun_out/ncu_full_d.log 2>&1; echo "ncu conv3 rc=$?"
$NCU -k regex:"border_gather|tokens_tc|tokens_tail|transformer_fwd|lidar_stem|pack_strip_kernel" -s 10 -c 5 -o gpurun_out/prof_chunk -f $PROF > gpurun_out/ncu_full_e.log 2>&1; echo "ncu chunk rc=$?"
if [ -n "$TRAIN" ]; then
TPROF="python bench.py --no-infer --no-graph --steps 1 --warmup 3"
$TPROF > gpurun_out/train_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/train_launches.csv $TPROF > gpurun_out/ncu_train_list.log 2>&1
echo "ncu train list rc=$?"
$TPROF > gpurun_out/train_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"wgrad_sps_tc|transformer_bwd|bn_bwd_apply" -s 18 -c 19 -o gpurun_out/prof_train -f $TPROF > gpurun_out/ncu_train_full.log 2>&1
echo "ncu train rc=$?"; tail -1 gpurun_out/ncu_train_full.log
fi
du -sh gpurun_out; ls -la gpurun_out | head -40
