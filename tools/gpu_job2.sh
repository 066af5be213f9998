#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for st in split linear encode layer model; do
  timeout 180 python tools/tc2_debug.py $st > gpurun_out/tc2_$st.log 2>&1; echo "rc=$?" >> gpurun_out/tc2_$st.log
done
tail -n 40 gpurun_out/tc2_*.log
