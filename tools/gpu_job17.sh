#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 60 python -u tools/tc2_debug.py layer > gpurun_out/tc2_layer.log 2>&1; grep -cE "^\[layer\]" gpurun_out/tc2_layer.log
GNB_TRACE=1 timeout 150 python tools/stress.py cfg3 120 > gpurun_out/stress.out 2> gpurun_out/stress.err; echo "stress cfg3 rc=$?"
grep -E "iter|STRESS" gpurun_out/stress.err | tail -2
grep "\[gnb\]" gpurun_out/stress.err | tail -2
GNB_TRACE=1 timeout 100 python tools/stress.py cfg2 800 > gpurun_out/stress2.out 2> gpurun_out/stress2.err; echo "stress cfg2 rc=$?"
grep -E "iter|STRESS" gpurun_out/stress2.err | tail -2
grep "\[gnb\]" gpurun_out/stress2.err | tail -2
