#!/bin/bash
# Round-2 ncu captures (run under gpurun, 1 GPU).  Each command first runs WITHOUT ncu and must exit 0 (profiling
# guide); then ONE launch of the named kernel is captured with --set full and exported as a raw CSV page.
#   gpurun --timeout 1500 -- 'bash tools/ncu_r2.sh'
mkdir -p gpurun_out
cap () {   # name, kernel regex, launches to skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > gpurun_out/r2_ncu_${name}_plain.log 2>&1 || { echo "$name: plain run failed"; tail -5 gpurun_out/r2_ncu_${name}_plain.log; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r2_prof_$name "$@" > gpurun_out/r2_ncu_${name}.log 2>&1
  ncu -i gpurun_out/r2_prof_$name.ncu-rep --page raw --csv > gpurun_out/r2_prof_${name}_raw.csv 2>/dev/null
  [ "$name" = "fused_c3" ] || rm -f gpurun_out/r2_prof_$name.ncu-rep      # gpurun_out/ may carry 64 MiB home
  python - "$name" <<'PY'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/r2_prof_{name}_raw.csv")))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
    print(name, {k: d.get(k) for k in keys})
PY
}
cap iso_acc    tv_accumulator_kernel 3 python tools/iso_bench.py --only two_pass --iters 3
cap iso_fused  tv_fused_iso_kernel   3 python tools/iso_bench.py --only fused --iters 3
cap fused_c3   tv_fused_kernel       3 python tools/plain_bench.py c3 --iters 3
cap fused_c1   tv_fused_kernel       3 python tools/plain_bench.py c1 --iters 3
cap fused_plain4d tv_fused_kernel    3 python tools/plain_bench.py plain4d --iters 3
cap fused_fp64 tv_fused_kernel       3 python tools/plain_bench.py fp64 --iters 3
ls -la gpurun_out/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
