#!/bin/bash
# bench lines of every workload (TAG in file names), then two ncu captures: k_extend in steady state (Cornell spp64, 4 M rays) and on mesh1m
TAG=${1:-v6}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for w in cornell book1 mesh1m book2 menger; do
  extra=""; [ $w = menger ] && extra="--steps 3"
  timeout 600 python bench.py --workload $w $extra > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w exit=$?"
done
timeout 900 python bench.py --workload mesh10m --spp 128 --steps 2 > gpurun_out/${TAG}_bench_mesh10m.json 2> gpurun_out/${TAG}_bench_mesh10m.err; echo "mesh10m exit=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_cornell.json 2>/dev/null; echo "reference exit=$?"
CMD="python bench.py --spp 64 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 8 -c 1 -o gpurun_out/${TAG}_extend_steady -f $CMD > gpurun_out/${TAG}_ncu_es.log 2>&1
CMDM="python bench.py --workload mesh1m --spp 8 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMDM > gpurun_out/${TAG}_plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 50 -c 2 -o gpurun_out/${TAG}_extend_mesh1m -f $CMDM > gpurun_out/${TAG}_ncu_m.log 2>&1
CMDX="python bench.py --workload mesh10m --spp 4 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMDX > gpurun_out/${TAG}_plain_x.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 3 -c 1 -o gpurun_out/${TAG}_extend_mesh10m -f $CMDX > gpurun_out/${TAG}_ncu_x.log 2>&1
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % os.environ.get("TAG","v6"))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}; c=d.get("cpu_baseline") or {}
        print(os.path.basename(f), "Mpaths/s %.1f Mrays/s %.1f ms/step %.1f e2e %.1f | frac %.2f ext %.2f shade %.2f | cpu %s" % (d["value"], d.get("mrays_per_s",0), d["ms_per_step"], d["e2e"]["value"], r.get("frac",0), r.get("extend_share_of_step",0), r.get("shade_share_of_step",0), c.get("value")))
    except Exception as e: print(f, "ERR", e)
PY
