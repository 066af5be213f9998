#!/bin/bash
# evict-first L2 hints on the e stream of the aggregation kernel: full GPU suite, short stress, cfg3 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -6
timeout 60 python tools/stress.py cfg3 20 > gpurun_out/stress.out 2> gpurun_out/stress.err; echo "stress cfg3 rc=$?"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cfg3.json'))
print('bench_cfg3', round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'])
for k,v in d['kernels'].items(): print('   ',k,round(v['ms_per_step']/v['launches_per_step'],2),'ms x',v['launches_per_step'])
PY
