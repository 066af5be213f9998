#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 90 python -u tools/tc2_debug.py layer model > gpurun_out/tc2_lm.log 2>&1; echo "rc=$?" >> gpurun_out/tc2_lm.log; grep -E "H=256|rc=|EXC|rror" gpurun_out/tc2_lm.log | tail -8
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python tools/edge_timing.py 256 > gpurun_out/edge_timing_256.log 2>&1; head -7 gpurun_out/edge_timing_256.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -2 gpurun_out/bench_cfg3.err
python - <<'PY'
import json
for f in ('bench_cfg3',):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as ex:
        print(f, 'no json', ex); continue
    print(f, round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'])
    for k,v in d['kernels'].items(): print('   ',k,round(v['ms_per_step']/v['launches_per_step'],2),'ms x',v['launches_per_step'])
PY
