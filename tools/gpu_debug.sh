#!/bin/bash
mkdir -p gpurun_out
echo "== default" > gpurun_out/debug.log
timeout 120 python tools/umma_diag.py c64_rand 64 11 5 300 rand >> gpurun_out/debug.log 2>&1
echo "== no cluster" >> gpurun_out/debug.log
HSV_UMMA_DEBUG=256 timeout 120 python tools/umma_diag.py c64_rand 64 11 5 300 rand >> gpurun_out/debug.log 2>&1
echo "== no cluster c128" >> gpurun_out/debug.log
HSV_UMMA_DEBUG=256 timeout 120 python tools/umma_diag.py c128_rand 128 7 1 300 rand >> gpurun_out/debug.log 2>&1
echo "== cluster c128 k3" >> gpurun_out/debug.log
timeout 120 python tools/umma_diag.py c128_k3 128 3 1 300 rand >> gpurun_out/debug.log 2>&1
echo "== sanitizer c128" >> gpurun_out/debug.log
timeout 300 compute-sanitizer --tool memcheck python tools/umma_diag.py c128_rand 128 7 1 300 rand 2>&1 | grep -v "^=========     " | head -60 >> gpurun_out/debug.log
cut -c1-400 gpurun_out/debug.log | tail -70
