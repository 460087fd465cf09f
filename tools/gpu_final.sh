#!/bin/bash
# final evidence of a build: GPU tests, smoke, bench lines of every workload, reference arm, launch list + ncu captures
TAG=${1:-v7}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in cornell book1 mesh1m book2 menger; do
  extra=""; [ $w = menger ] && extra="--steps 3"
  timeout 600 python bench.py --workload $w $extra > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w exit=$?"
done
timeout 900 python bench.py --workload mesh10m --spp 128 --steps 2 > gpurun_out/${TAG}_bench_mesh10m.json 2> gpurun_out/${TAG}_bench_mesh10m.err; echo "mesh10m exit=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_cornell.json 2>/dev/null; echo "reference exit=$?"
CMD="python bench.py --spp 16 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 30 -c 2 -o gpurun_out/${TAG}_extend -f $CMD > gpurun_out/${TAG}_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 30 -c 2 -o gpurun_out/${TAG}_shade -f $CMD > gpurun_out/${TAG}_ncu_s.log 2>&1
CMDM="python tools/gpu_one_render.py mesh1m 8"
$CMDM > gpurun_out/${TAG}_plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 2 -c 2 -o gpurun_out/${TAG}_extend_mesh1m -f $CMDM > gpurun_out/${TAG}_ncu_m.log 2>&1
CMDX="python tools/gpu_one_render.py mesh10m 2"
$CMDX > gpurun_out/${TAG}_plain_x.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 3 -c 1 -o gpurun_out/${TAG}_extend_mesh10m -f $CMDX > gpurun_out/${TAG}_ncu_x.log 2>&1
TAG=$TAG python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % os.environ["TAG"])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}; c=d.get("cpu_baseline") or {}
        print(os.path.basename(f), "Mpaths/s %.1f Mrays/s %.1f ms/step %.1f e2e %.1f | frac %.2f ext %.2f shade %.2f | cpu %s" % (d["value"], d.get("mrays_per_s",0), d["ms_per_step"], d["e2e"]["value"], r.get("frac",0), r.get("extend_share_of_step",0), r.get("shade_share_of_step",0), c.get("value")))
    except Exception as e: print(f, "ERR", e)
PY
