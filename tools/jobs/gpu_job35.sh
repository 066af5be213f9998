#!/bin/bash
# head-of-branch verification: smoke, full GPU suite, cfg3 / cfg2 bench, reference arm, ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 360 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.err
timeout 120 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-260 gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
gzip -f gpurun_out/launches_cfg3.csv
python - <<'PY'
import json
for f in ('bench_cfg3','bench_cfg2'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as ex:
        print(f, 'no json', ex); continue
    print(f, round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'], 'launches', d.get('gpu_launches'))
    for k,v in d['kernels'].items(): print('   ',k,round(v['ms_per_step']/v['launches_per_step'],2),'ms x',v['launches_per_step'])
    print('   ', d.get('cpu_baseline'))
PY
