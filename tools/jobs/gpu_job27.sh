#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"score_forward_tc2" -s 1 -c 1 -o gpurun_out/prof4_score python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out/prof4_score.ncu-rep
