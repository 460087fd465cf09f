#!/bin/bash
# multi-GPU validation in one gpurun call (run with gpurun --gpus 2): tests of the in-process multi-device handle, then bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/multi_tests.log
python bench.py --no-per-config --steps 3 > gpurun_out/multi_bench_n1.json 2> gpurun_out/multi_bench_n1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-per-config --steps 3 > gpurun_out/multi_bench_n2.json 2> gpurun_out/multi_bench_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/multi_bench_n1.err gpurun_out/multi_bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/multi_bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"], d.get("weak"))
    except Exception as e:
        print(n, "no line:", e)
PY
