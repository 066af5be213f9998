#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/tc2_debug.py layers > gpurun_out/tc2_det.log 2>&1; echo "rc=$?" >> gpurun_out/tc2_det.log
cat gpurun_out/tc2_det.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
