#!/bin/bash
# Round evidence for the inference path: gpu tests, bench (both arms), ncu launch list and
# full captures of the two dominant kernels (B200_PROFILING.md recipe).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
PROF="python bench.py --steps 1 --warmup 3 --windows 16384 --no-cpu --no-e2e"
$PROF > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log
$PROF > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_sps_tc -s 12 -c 6 -o gpurun_out/prof_conv $PROF > gpurun_out/ncu_full.log 2>&1
echo "ncu conv rc=$?"; tail -2 gpurun_out/ncu_full.log
$PROF > gpurun_out/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"transformer_fwd|pack_sps" -s 6 -c 3 -o gpurun_out/prof_tokens $PROF > gpurun_out/ncu_tokens.log 2>&1
echo "ncu tokens rc=$?"; tail -2 gpurun_out/ncu_tokens.log
ls -la gpurun_out
