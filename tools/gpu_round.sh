#!/bin/bash
# full GPU test suite, smoke, token-kernel timing at P = 9 / 10 / 11, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
for p in 9 10 11; do for k in tc tm3 tm4; do
  P=$p VITCNN_TC_KERNEL=$k ONLY=tcgen05 N=131072 timeout 120 python tools/time_tokens.py 2>&1 | tail -1 | tee -a gpurun_out/tm_time_p.log
done; done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_b7.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d['roofline']['limiter']['frac_of_floor'])
P
