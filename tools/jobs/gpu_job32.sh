#!/bin/bash
# BatchNorm affine passes centred on the column mean: training tests, hand-off tests, gradient check vs fp64
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_handoff.py -q > gpurun_out/pytest_train.log 2>&1; echo "rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_train.log | tail -15
for k in sym rev; do
  timeout 200 python tests/diag/grad_check.py $k > gpurun_out/grad_check_$k.txt 2>&1
  head -2 gpurun_out/grad_check_$k.txt; echo "flagged: $(grep -c '<--' gpurun_out/grad_check_$k.txt)"; sort -k2 -g -r gpurun_out/grad_check_$k.txt | grep -v "e+0\|nan.*1.00e-12" | head -4
done
timeout 200 python tests/diag/grad_dump.py > /dev/null 2>&1
