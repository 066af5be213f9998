#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_forward_tc2|node_update2" -s 18 -c 2 -o gpurun_out/prof2_mid python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
