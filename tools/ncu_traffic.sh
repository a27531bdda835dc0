#!/bin/bash
# DRAM bytes, duration and cache hit rates of ONE launch of a kernel inside bench.py (run under gpurun, 1 GPU).
#   KREGEX=tv_fused SKIP=3 BENCH_EXTRA="--schedule fused" bash tools/ncu_traffic.sh
# The same command is first run without ncu and must exit 0 (profiling guide).
mkdir -p gpurun_out
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu ${BENCH_EXTRA}"
$CMD > gpurun_out/ncu_traffic_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_traffic_plain.log; exit 1; }
ncu $M --clock-control none -k regex:${KREGEX:-tv_fused} -s ${SKIP:-3} -c 1 --csv --log-file gpurun_out/ncu_traffic.csv $CMD > /dev/null 2>&1
grep -E "dram__|gpu__time|hit_rate" gpurun_out/ncu_traffic.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
