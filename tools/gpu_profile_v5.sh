#!/bin/bash
# v5 evidence: plain bench run, launch list, ncu --set full of k_extend / k_shade (Cornell spp16) and k_extend (mesh1m spp8), mesh10m bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --spp 16 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/v5_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/v5_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/v5_launches.csv $CMD > gpurun_out/v5_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 30 -c 2 -o gpurun_out/v5_extend -f $CMD > gpurun_out/v5_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 30 -c 2 -o gpurun_out/v5_shade -f $CMD > gpurun_out/v5_ncu_s.log 2>&1
CMDM="python bench.py --workload mesh1m --spp 8 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMDM > gpurun_out/v5_plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 12 -c 2 -o gpurun_out/v5_extend_mesh1m -f $CMDM > gpurun_out/v5_ncu_m.log 2>&1
timeout 900 python bench.py --workload mesh10m --spp 128 --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/v5_mesh10m.json 2> gpurun_out/v5_mesh10m.err
tail -c 600 gpurun_out/v5_mesh10m.json
