#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"score_forward_tc2|encode2" -s 3 -c 3 -o gpurun_out/prof3_mid python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/prof3_mid.ncu-rep
