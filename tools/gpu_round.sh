#!/bin/bash
# one GPU round: parity tests, then the option sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log
timeout 900 python tools/gpu_sweep.py "$@" 2>&1 | tee gpurun_out/sweep.log
