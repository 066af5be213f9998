#!/bin/bash
# ncu --set full of the aggregation kernel (where do the DRAM reads above the algorithmic bytes come from?) and of the
# score kernel with the lane = unit phase B, both at the `mid` workload (2M / 12M, H = 256)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"edge_forward_tc2" -s 9 -c 1 -o gpurun_out/prof5_edge python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"score_forward_tc2" -s 1 -c 1 -o gpurun_out/prof5_score python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full5b.log 2>&1
ls -la gpurun_out/prof5_*.ncu-rep
