#!/bin/bash
# hand-off + subgraph tests, timing of the hand-off kernels at cfg3 size, full GPU suite, cfg3 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_handoff.py -q > gpurun_out/pytest_handoff.log 2>&1; echo "handoff rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_handoff.log | tail -15
timeout 200 python tools/handoff_timing.py > gpurun_out/handoff_timing.txt 2>&1; cat gpurun_out/handoff_timing.txt | tail -5
timeout 360 python -m pytest tests -m gpu -q --deselect tests/test_gpu_handoff.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.err
python - <<'PY'
import json
for f in ('bench_cfg3',):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as ex:
        print(f, 'no json', ex); continue
    print(f, round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,1), 'M edges/s e2e', round(d['e2e']['value']/1e6,1), 'roofline', round(d['roofline']['frac'],3), d['clocks'])
PY
