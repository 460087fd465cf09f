#!/bin/bash
# final evidence of a build: GPU tests, smoke, the driver's bench line (headline + per_config), reference arm, launch list + ncu captures
TAG=${1:-r02_v1}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench_mesh1m.json 2> gpurun_out/${TAG}_bench_mesh1m.err; echo "bench exit=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo "reference exit=$?"
# launch list of the same command as the bench line (shares of the step), then one steady-state launch of k_extend per workload
CMD="python bench.py --steps 2 --warmup 3 --spp 32 --no-cpu-baseline --no-per-config --no-parity"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_mesh1m_spp32.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
export MRT_POOL_SLOTS=16777216  # the captures (and tools/ncu_side_files.py) are per launch of 2^24 rays
for spec in "cornell 32 8" "mesh1m 32 5" "mesh10m 4 4" "book2 32 3"; do
  set -- $spec
  C="python tools/gpu_one_render.py $1 $2"
  $C > gpurun_out/${TAG}_plain_$1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend --launch-skip $3 --launch-count 1 -o gpurun_out/${TAG}_extend_$1 -f $C > gpurun_out/${TAG}_ncu_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_extend_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_extend_$1_raw.csv 2>/dev/null
done
C="python tools/gpu_one_render.py cornell 32"
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 8 --launch-count 1 -o gpurun_out/${TAG}_shade_cornell -f $C > gpurun_out/${TAG}_ncu_s.log 2>&1
ncu -i gpurun_out/${TAG}_shade_cornell.ncu-rep --page raw --csv > gpurun_out/${TAG}_shade_cornell_raw.csv 2>/dev/null
TAG=$TAG python - <<'PY'
import json,glob,os
TAG=os.environ.get("TAG","")
for f in sorted(glob.glob("gpurun_out/*_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}; c=d.get("cpu_baseline") or {}
        print(os.path.basename(f), "Mpaths/s %.1f Mrays/s %.1f ms/step %.1f e2e %.1f | frac %.2f ext %.2f shade %.2f | cpu %s" % (d["value"], d.get("mrays_per_s",0), d["ms_per_step"], d["e2e"]["value"], r.get("frac",0), r.get("extend_share_of_step",0), r.get("shade_share_of_step",0), c.get("value")))
        for k,v in (d.get("per_config") or {}).items(): print("   ", k, "Mpaths/s %.1f e2e %.1f ms/step %.2f" % (v["value"], v["e2e"]["value"], v["ms_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
