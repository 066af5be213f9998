#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -u tests/diag/tc2_debug.py layer > gpurun_out/memcheck_layer.log 2>&1; echo "memcheck layer rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|\[layer\]" gpurun_out/memcheck_layer.log | tail -12
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shipped_model or other_models or empty or score_tc2 or encode or split16" > gpurun_out/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/memcheck_tests.log | tail -8
