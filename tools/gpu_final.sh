mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
PROF="python bench.py --steps 1 --warmup 3 --windows 65536 --no-cpu --no-e2e --no-train"
$PROF > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none -k regex:"border_gather|tokens_tc|tokens_tail|transformer_fwd|lidar_stem|pack_strip_kernel" -s 10 -c 5 -o gpurun_out/prof_chunk -f $PROF > gpurun_out/ncu_full_e.log 2>&1; echo "chunk rc=$?"
timeout 600 python tools/bench_shapes.py 2>&1 | grep "^{" > gpurun_out/shapes.jsonl; echo "shapes rc=$?"
