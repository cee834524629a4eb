#!/bin/bash
# compute-sanitizer over a small end-to-end workload; summaries land in gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?" > gpurun_out/sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver ok|Error:|Invalid|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -30 >> gpurun_out/sanitizer_$tool.txt
  cat gpurun_out/sanitizer_$tool.txt
done
