# Round-end evidence at the default settings (chunk 131072, 6 block-aligned sub-bands): gpu tests, smoke, the bench
# line, the chunk sweep with its bit-exactness check, the ncu launch list of one profiled scene and one full capture
# of the token kernel (dram bytes per launch for roofline.traffic).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 200 python tools/chunk_sweep.py 32768 65536 131072 196608 262144 > gpurun_out/chunk_sweep.jsonl 2> gpurun_out/chunk_sweep.err; echo "sweep rc=$?"
PROF="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-train"
timeout 100 $PROF > gpurun_out/prof_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:tokens_tc_kernel -s 6 -c 1 -o gpurun_out/prof_tokens_c131072 -f $PROF > gpurun_out/ncu_full_tokens.log 2>&1; echo "tokens rc=$?"
ncu -i gpurun_out/prof_tokens_c131072.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size > gpurun_out/tokens_c131072_raw.csv 2>&1
du -sh gpurun_out
