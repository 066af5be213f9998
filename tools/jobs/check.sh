#!/bin/bash
# GPU verification pass, run as  gpurun [--gpus N] -- 'bash tools/jobs/check.sh [steps...]'  from the repository root.
# Steps (default: "smoke tests bench"):
#   smoke            __graft_entry__.smoke()
#   tests[:EXPR]     pytest -m gpu (optionally -k EXPR)
#   bench[:WL]       bench.py --workload WL (default cfg3) -> gpurun_out/bench_WL.json + a one-line summary
#   benchq[:WL]      the same without the CPU legs (kernel work only; -> gpurun_out/benchq_WL.json)
#   edgetime[:H]     tools/edge_timing.py: cycle accounting of the aggregation kernel's epilogue warps
#   modes[:WL:REPS]  tools/edge_modes.py: launch time of the aggregation kernel under each L2 management setting
#   ref              bench.py --impl reference
#   stress[:WL:N:P]  P fresh processes x N forwards of workload WL (tools/stress.py), default cfg3:100:3
#   dist:N[:WL]      tools/dist_check.py (inference + training equivalence) + bench.py --workload WL (default cfg3) under torchrun on N GPUs
#   distbench:N:WL[:TAG]  only the bench of WL under torchrun on N GPUs (-> gpurun_out/bench_WL_xN[TAG].json)
#   sanitize[:TOOL]  compute-sanitizer --tool TOOL (default synccheck) over the small liveness / parity tests
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
steps=("$@"); [ ${#steps[@]} -eq 0 ] && steps=(smoke tests bench)
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as exc:
    print('no bench line:', exc); sys.exit(0)
if d.get('error'):
    print('BENCH ERROR', d['error'], d.get('hang_report')); sys.exit(0)
r, e = d.get('roofline') or {}, d.get('e2e') or {}
print(sys.argv[1], round(d['ms_per_step'], 1), 'ms', round(d['value'] / 1e6, 1), 'M edges/s  e2e', round((e.get('value') or 0) / 1e6, 1),
      ' roofline', round(r.get('frac') or 0, 3), round(r.get('ms_per_launch') or 0, 2), 'ms/launch', d.get('clocks'))
for k, v in (d.get('kernels') or {}).items():
    print('    %-28s %8.2f ms x %g' % (k, v['ms_per_step'] / max(v['launches_per_step'], 1e-9), v['launches_per_step']))
c = d.get('cpu_baseline')
if c: print('    cpu', round(c['value']), 'edges/s on', c['cores'], 'cores; parity', c.get('parity_max_prob_err_on_sample'))
PY
}
for st in "${steps[@]}"; do
  IFS=: read -r name a1 a2 a3 <<< "$st"
  case "$name" in
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    tests) timeout 1500 python -m pytest tests -m gpu -q -x ${a1:+-k "$a1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
           grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -8 ;;
    bench) wl=${a1:-cfg3}; timeout 900 python bench.py --workload "$wl" --steps 5 --warmup 3 > "gpurun_out/bench_$wl.json" 2> "gpurun_out/bench_$wl.err"
           echo "bench $wl rc=$?"; summ "gpurun_out/bench_$wl.json" ;;
    benchq) wl=${a1:-cfg3}; timeout 600 python bench.py --workload "$wl" --steps 5 --warmup 3 --no-cpu > "gpurun_out/benchq_$wl.json" 2> "gpurun_out/benchq_$wl.err"
           echo "benchq $wl rc=$?"; summ "gpurun_out/benchq_$wl.json" ;;
    edgetime) timeout 300 python tools/edge_timing.py ${a1:-256} > "gpurun_out/edge_timing_${a1:-256}.txt" 2>&1; grep -v "CTA 0" "gpurun_out/edge_timing_${a1:-256}.txt" | tail -9 ;;
    modes) timeout 600 python tools/edge_modes.py ${a1:-cfg3} ${a2:-24} ${a3:-} > "gpurun_out/edge_modes_${a1:-cfg3}.txt" 2>&1; tail -10 "gpurun_out/edge_modes_${a1:-cfg3}.txt" ;;
    power) timeout 600 python tools/power_probe.py ${a1:-cfg3} ${a2:-3} > "gpurun_out/power_${a1:-cfg3}.txt" 2>&1; tail -5 "gpurun_out/power_${a1:-cfg3}.txt" ;;
    trywait) timeout 120 tools/microbench/trywait_probe > gpurun_out/trywait_probe.txt 2>&1; cat gpurun_out/trywait_probe.txt ;;
    ref)   timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json ;;
    stress) wl=${a1:-cfg3}; n=${a2:-100}; p=${a3:-3}
           for i in $(seq 1 "$p"); do timeout 900 python tools/stress.py "$wl" "$n" > /dev/null 2> "gpurun_out/stress_$i.err"; echo "stress $wl x$n process $i rc=$? $(grep -c ' ok ' gpurun_out/stress_$i.err) forwards ok; $(tail -1 gpurun_out/stress_$i.err)"; done ;;
    dist)  n=${a1:-2}
           timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > "gpurun_out/dist_check_x$n.log" 2>&1; echo "dist_check x$n rc=$?"; tail -4 "gpurun_out/dist_check_x$n.log"
           wl=${a2:-cfg3}
           timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$n" --workload "$wl" --steps 5 --warmup 3 > "gpurun_out/bench_${wl}_x$n.json" 2> "gpurun_out/bench_${wl}_x$n.err"; echo "bench $wl x$n rc=$?"; summ "gpurun_out/bench_${wl}_x$n.json" ;;
    distbench) n=${a1:-2}; wl=${a2:-cfg3}; tag=${a3:-}
           timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus "$n" --workload "$wl" --steps 5 --warmup 3 > "gpurun_out/bench_${wl}_x$n$tag.json" 2> "gpurun_out/bench_${wl}_x$n$tag.err"; echo "bench $wl x$n $tag rc=$?"; summ "gpurun_out/bench_${wl}_x$n$tag.json" ;;
    sanitize) tool=${a1:-synccheck}
           timeout 1500 compute-sanitizer --tool "$tool" python -m pytest tests/test_gpu_liveness.py -q -x -k "stalled and 3000" > "gpurun_out/sanitize_$tool.log" 2>&1; echo "sanitize $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" "gpurun_out/sanitize_$tool.log" | tail -3 ;;
    *) echo "unknown step $st" ;;
  esac
done
