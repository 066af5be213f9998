#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
tail -40 gpurun_out/pytest_train.log
