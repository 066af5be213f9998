#!/bin/bash
# one gpurun call: GPU tests, bench at cfg3/cfg2, launch list + full ncu capture at the profiling size
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
gzip -f gpurun_out/launches_mid.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_forward_tc|node_update|node_linear_tc|score_forward|encode_kernel" -s 28 -c 5 -o gpurun_out/prof_mid python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"score_forward" -s 1 -c 1 -o gpurun_out/prof_mid_score python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full2.log 2>&1
du -sh gpurun_out; ls -la gpurun_out
