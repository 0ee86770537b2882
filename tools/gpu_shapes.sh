#!/bin/bash
# BASELINE.json configs[3] / [4] on one GPU: other raster shapes and patch sizes through bench.py
mkdir -p gpurun_out
for cfg in "muufl 11" "augsburg 11" "houston 7" "houston 9" "houston 15"; do
  set -- $cfg
  timeout 400 python bench.py --workload $1 --patch $2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_shape_$1_$2.json 2> gpurun_out/r02_shape_$1_$2.err
  python - "$1" "$2" <<'P'
import json, sys
d = json.loads(open(f"gpurun_out/r02_shape_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], sys.argv[2], round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), round(d["train"]["value"] / 1e6, 3))
P
done
