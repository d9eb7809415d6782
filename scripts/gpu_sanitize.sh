#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok|hazard|Error" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
