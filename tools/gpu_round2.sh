#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
for w in mesh1m; do
  timeout 600 python bench.py --workload $w > gpurun_out/v7_bench_$w.json 2> gpurun_out/v7_bench_$w.err; echo "$w exit=$?"
  timeout 600 python bench.py --workload $w --bvh sah --no-cpu-baseline > gpurun_out/v7_bench_${w}_sah.json 2> gpurun_out/v7_bench_${w}_sah.err; echo "$w sah exit=$?"
done
timeout 900 python bench.py --workload mesh10m --spp 128 --steps 2 > gpurun_out/v7_bench_mesh10m.json 2> gpurun_out/v7_bench_mesh10m.err; echo "mesh10m exit=$?"
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/v7_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(os.path.basename(f), "Mpaths/s %.1f Mrays/s %.1f ms/step %.1f e2e %.1f (%.1f ms) | nodes/ray %.2f" % (d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["per_ray"]["node_visits"]))
    except Exception as e: print(f, "ERR", e)
PY
