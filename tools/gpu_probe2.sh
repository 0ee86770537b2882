#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" | tail -1
for m in 0 1 2 3; do echo "=== umma probe mode $m"; timeout 120 ./tools/umma_probe.bin 148 $m 2>&1 | tail -60; done > gpurun_out/umma_probe.txt 2>&1
cat gpurun_out/umma_probe.txt | head -150
for c in 4096 8192 16384; do echo "=== chunk $c"; python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --chunk $c 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'])
"; done
PROF="python bench.py --steps 1 --warmup 3 --windows 4096 --no-cpu --no-e2e"
$PROF > gpurun_out/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:transformer_fwd -s 3 -c 1 -o gpurun_out/prof_tokens $PROF > gpurun_out/ncu_tokens.log 2>&1
echo "ncu tokens rc=$?"
