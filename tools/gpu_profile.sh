#!/bin/bash
# ncu evidence for the current kernels: launch list + --set full of k_extend / k_shade (Cornell spp16) and k_extend (mesh1m spp8)
TAG=${1:-v6}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --spp 16 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 30 -c 2 -o gpurun_out/${TAG}_extend -f $CMD > gpurun_out/${TAG}_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 30 -c 2 -o gpurun_out/${TAG}_shade -f $CMD > gpurun_out/${TAG}_ncu_s.log 2>&1
CMDM="python bench.py --workload mesh1m --spp 8 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMDM > gpurun_out/${TAG}_plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_extend -s 35 -c 2 -o gpurun_out/${TAG}_extend_mesh1m -f $CMDM > gpurun_out/${TAG}_ncu_m.log 2>&1
tail -c 300 gpurun_out/${TAG}_plain.log
