#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check.log
tail -12 gpurun_out/dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_cfg3_x2.json 2> gpurun_out/bench_cfg3_x2.err; echo "rc=$?" >> gpurun_out/bench_cfg3_x2.err
tail -5 gpurun_out/bench_cfg3_x2.err; cat gpurun_out/bench_cfg3_x2.json | cut -c1-1500
