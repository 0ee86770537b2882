# A/B of the host pipeline's sub-band cuts (scene.py): equal parts vs whole block rows at 8 / 6 / 5 sub-bands,
# then the full default bench line with the best setting.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_scene.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-train"
VITCNN_SUBBAND_SPLIT=equal timeout 120 $B > gpurun_out/ab_equal8.json 2> gpurun_out/ab_equal8.err
for n in 8 7 6 5; do VITCNN_PIPELINE=$n timeout 120 $B > gpurun_out/ab_blocks$n.json 2> gpurun_out/ab_blocks$n.err; done
for n in 6 5; do VITCNN_SUBBAND_LEAD=even VITCNN_PIPELINE=$n timeout 120 $B > gpurun_out/ab_even$n.json 2> gpurun_out/ab_even$n.err; done
python - <<'PY'
import json, glob
best = None
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        ms = j["e2e"]["ms_per_step"]
        print(f, "device ms", round(j["ms_per_step"], 2), "e2e ms", round(ms, 2))
        if "blocks" in f and (best is None or ms < best[0]):
            best = (ms, f.split("blocks")[1].split(".")[0])
    except Exception as e:
        print(f, "unreadable", e)
open("gpurun_out/best_pipeline", "w").write(best[1] if best else "8")
PY
export VITCNN_PIPELINE=$(cat gpurun_out/best_pipeline); echo "best pipeline $VITCNN_PIPELINE"
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
