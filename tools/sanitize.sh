#!/bin/bash
# compute-sanitizer memcheck + racecheck over tools/sanitizer_cases.py (run under gpurun, 1 GPU):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# The plain run must pass first.  Logs: gpurun_out/r2_sanitizer_{memcheck,racecheck}.log (copied to profiles/).
mkdir -p gpurun_out
timeout 300 python tools/sanitizer_cases.py > gpurun_out/r2_sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r2_sanitizer_plain.log; exit 1; }
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_cases.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer cases ok|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -12
  tail -4 gpurun_out/r2_sanitizer_$tool.log | cut -c1-300
done
