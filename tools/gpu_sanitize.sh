#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|error" gpurun_out/sanitize_$tool.log | head -8
done
