#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/edge_timing.py 256 > gpurun_out/edge_timing_256.log 2>&1
timeout 300 python tools/edge_timing.py 128 > gpurun_out/edge_timing_128.log 2>&1
cat gpurun_out/edge_timing_256.log gpurun_out/edge_timing_128.log
