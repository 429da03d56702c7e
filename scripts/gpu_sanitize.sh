#!/bin/bash
# compute-sanitizer over every kernel (small shapes): memcheck, racecheck, synccheck.
set +e
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_targets.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize targets done" gpurun_out/sanitize_$tool.log | head -8
done
