#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
GNB_TRACE=1 timeout 120 python tools/stress.py cfg2s 600 > gpurun_out/stress3.out 2> gpurun_out/stress3.err; echo "stress cfg2s rc=$?"
grep -E "iter|STRESS" gpurun_out/stress3.err | tail -2
GNB_TRACE=1 timeout 120 python tools/stress.py small 1500 > gpurun_out/stress4.out 2> gpurun_out/stress4.err; echo "stress small rc=$?"
grep -E "iter|STRESS" gpurun_out/stress4.err | tail -2
