#!/bin/bash
# Round-2 evidence (B200_PROFILING.md recipe): ncu launch lists of one inference step and one training step, full captures of
# the kernels that are new this round.  Every ncu command runs only after the same command has exited 0 without ncu.
mkdir -p gpurun_out
PROF="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-train"
$PROF > gpurun_out/ev_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ev_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches.csv $PROF > gpurun_out/ev_ncu_list.log 2>&1
echo "ncu list rc=$?"
NCU="ncu --set full --clock-control none"
$NCU -k regex:conv_var_kernel -s 12 -c 4 -o gpurun_out/ev_conv_var -f $PROF > gpurun_out/ev_ncu_a.log 2>&1; echo "ncu conv_var rc=$?"
$NCU -k regex:conv_sps_tc2_kernel -s 31 -c 1 -o gpurun_out/ev_conv1 -f $PROF > gpurun_out/ev_ncu_b.log 2>&1; echo "ncu conv1 rc=$?"
GP="python tools/time_gather.py"
$GP > gpurun_out/ev_gather_plain.log 2>&1 && $NCU -k regex:gather_tma_kernel -s 3 -c 1 -o gpurun_out/ev_gather -f $GP > gpurun_out/ev_ncu_c.log 2>&1; echo "ncu gather rc=$?"
TPROF="python bench.py --no-infer --no-graph --no-cpu --steps 1 --warmup 3"
$TPROF > gpurun_out/ev_train_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/ev_train_launches.csv $TPROF > gpurun_out/ev_ncu_train.log 2>&1
echo "ncu train list rc=$?"
du -sh gpurun_out; ls -la gpurun_out | grep ev_
