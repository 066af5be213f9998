#!/bin/bash
# hand-off tests + score kernel with lane = unit phase B: full GPU test suite, cfg3 / cfg2 bench, FFMA2 micro-benchmark
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 360 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 60 tools/microbench/ffma2_bench > gpurun_out/ffma2_bench.txt 2>&1; cat gpurun_out/ffma2_bench.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.err
timeout 120 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python - <<'PY'
import json
for f in ('bench_cfg3','bench_cfg2'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as ex:
        print(f, 'no json', ex); continue
    print(f, round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'])
    for k,v in d['kernels'].items(): print('   ',k,round(v['ms_per_step']/v['launches_per_step'],2),'ms x',v['launches_per_step'])
PY
