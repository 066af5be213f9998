#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/tc2_debug.py model > gpurun_out/tc2_model.log 2>&1; echo "rc=$?" >> gpurun_out/tc2_model.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
cat gpurun_out/tc2_model.log
