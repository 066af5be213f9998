#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_x$N.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_x$N.log
grep -E "dist x|DIST|rc=|Error" gpurun_out/dist_check_x$N.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_cfg3_x$N.json 2> gpurun_out/bench_cfg3_x$N.err; echo "rc=$?" >> gpurun_out/bench_cfg3_x$N.err
tail -3 gpurun_out/bench_cfg3_x$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_x$N.json'))
print('x$N', round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'])
for k,v in d['kernels'].items(): print('   ',k,round(v['ms_per_step']/v['launches_per_step'],2),'ms x',v['launches_per_step'])
PY
