#!/bin/bash
# bench + ncu launch list + one full capture of the dominant kernel (B200_PROFILING.md recipe)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
PROF="python bench.py --steps 1 --warmup 3 --windows 4096 --no-cpu --no-e2e"
$PROF > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log
$PROF > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_sps_tc -s 6 -c 3 -o gpurun_out/prof_conv $PROF > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
