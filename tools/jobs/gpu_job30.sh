#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for k in bce rev; do
  timeout 200 python tests/diag/grad_check.py $k > gpurun_out/grad_check_$k.txt 2>&1
  head -5 gpurun_out/grad_check_$k.txt; grep "convs.7\|predictor" gpurun_out/grad_check_$k.txt | grep -v "bias"
done
