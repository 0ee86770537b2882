#!/bin/bash
# final N = 1 record of the round: bench line, CPU arm, ncu launch list, full capture of tokens_tm_kernel on the scene path
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/tm_bench_n1.json 2> gpurun_out/tm_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/tm_bench_ref_n1.json 2> gpurun_out/tm_bench_ref_n1.err; echo "ref rc=$?"
PROF="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-train"
timeout 200 $PROF > gpurun_out/tm_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/tm_plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/tm_launches.csv $PROF > gpurun_out/tm_ncu_list.log 2>&1
echo "ncu list rc=$?"
PROFW="$PROF --windows 131072"     # one token launch of 131 072 windows per step
timeout 200 $PROFW > gpurun_out/tm_plain_w.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tokens_tm_kernel -s 3 -c 1 -o gpurun_out/tm_tokens_scene -f $PROFW > gpurun_out/tm_ncu_full.log 2>&1
echo "ncu tokens rc=$?"
