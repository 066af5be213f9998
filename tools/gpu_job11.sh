#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for st in linear layer model; do
  timeout 60 python -u tools/tc2_debug.py $st > gpurun_out/tc2_$st.log 2>&1; echo "rc=$?" >> gpurun_out/tc2_$st.log
  tail -12 gpurun_out/tc2_$st.log
done
