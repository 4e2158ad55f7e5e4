#!/bin/bash
# full GPU test-suite + compute-sanitizer (memcheck / racecheck / synccheck) at the current HEAD
mkdir -p gpurun_out
P=gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu -s > ${P}_t_all.log 2>&1; echo "tests rc=$?"; tail -2 ${P}_t_all.log; grep "\[parity\]" ${P}_t_all.log > ${P}_parity.log
for tool in memcheck racecheck synccheck; do
  EXTRA=""; [ $tool = racecheck ] && EXTRA="--kernel-regex-exclude kns=mha_mma"
  timeout 900 compute-sanitizer --tool $tool $EXTRA --print-limit 20 python tools/sanitize_target.py > ${P}_sanitize_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" ${P}_sanitize_$tool.log | head -4
done
