#!/bin/bash
# launch list + full ncu capture of the tc2 kernels at the profiling size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_mid.csv python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
gzip -f gpurun_out/launches_mid.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_forward_tc2|node_update2|node_linear_tc2|score_forward_tc2|encode2" -s 30 -c 5 -o gpurun_out/prof_mid python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"score_forward_tc2" -s 1 -c 1 -o gpurun_out/prof_mid_score python bench.py --workload mid --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full2.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | head -30
