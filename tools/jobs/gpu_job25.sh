#!/bin/bash
# bench.py under torchrun on all GPUs of the box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_cfg3_x$N.json 2> gpurun_out/bench_cfg3_x$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_cfg3_x$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_x$N.json'))
print('x$N', round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'])
tot=0
for k,v in d['kernels'].items(): print('   ',k,round(v['ms_per_step']/v['launches_per_step'],2),'ms x',v['launches_per_step']); tot+=v['ms_per_step']
print('sum of kernel ms', round(tot,1), 'of', round(d['ms_per_step'],1))
PY
