#!/bin/bash
# debugging aid: DRAM bytes / L2 hit rate of the fused kernel for several CYTVDN_FUSED_HINT values
mkdir -p gpurun_out
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu ${BENCH_EXTRA}"
$CMD > /dev/null 2>&1 || exit 1
for h in ${HINTS:-0}; do
  CYTVDN_FUSED_HINT=$h ncu $M --clock-control none -k regex:${KREGEX:-tv_fused} -s ${SKIP:-3} -c 1 --csv --log-file gpurun_out/dbg_$h.csv $CMD > /dev/null 2>&1
  echo "== hint $h"; grep -E "dram__|gpu__time|hit_rate" gpurun_out/dbg_$h.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
