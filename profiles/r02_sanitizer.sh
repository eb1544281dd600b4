#!/bin/bash
# compute-sanitizer over the small configurations (SURVEY §5 "race detection"): memcheck, racecheck (shared memory hazards), initcheck
# (uninitialised global memory reads), synccheck. Logs: gpurun_out/r02_sanitizer_<tool>.log; the last lines hold the error summary.
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize.py 64 48 > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done|Error|hazard" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | head -12
done
