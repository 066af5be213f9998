#!/bin/bash
# ncu captures, run as  gpurun -- 'bash tools/jobs/profile.sh WORKLOAD [launches|full:REGEX ...]'  (one GPU only).
#   launches      every launch of one forward with its device time -> gpurun_out/launches_WL.csv
#   metrics:REGEX the DRAM traffic / instruction / pipe counters of 2 launches of the kernels matching REGEX (a few replay
#                 passes instead of the ~40 of --set full: minutes instead of a quarter of an hour at cfg3) -> gpurun_out/metrics_WL_REGEX.csv
#   full:REGEX    ncu --set full of the kernels matching REGEX (3 launches after the warm-up ones) -> gpurun_out/full_WL_REGEX.ncu-rep
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
wl=${1:-mid}; shift
for st in "$@"; do
  IFS=: read -r name a1 <<< "$st"
  case "$name" in
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "gpurun_out/launches_$wl.csv" \
                python bench.py --workload "$wl" --steps 2 --warmup 1 --no-cpu > /dev/null 2> "gpurun_out/launches_$wl.err"; echo "launches rc=$?" ;;
    full)     timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$a1" -s 6 -c 3 -f -o "gpurun_out/full_${wl}_$a1" \
                python tools/stress.py "$wl" 2 > /dev/null 2> "gpurun_out/full_${wl}_$a1.err"; echo "full $a1 rc=$?" ;;
    metrics)  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max \
                --clock-control none -k "regex:$a1" -s 8 -c 2 --csv --log-file "gpurun_out/metrics_${wl}_$a1.csv" \
                python tools/stress.py "$wl" 2 > /dev/null 2> "gpurun_out/metrics_${wl}_$a1.err"; echo "metrics $a1 rc=$?" ;;
    *) echo "unknown step $st" ;;
  esac
done
