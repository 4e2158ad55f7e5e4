#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
for dbg in 0 32; do
echo "=== HSV_UMMA_DEBUG=$dbg"
HSV_UMMA_DEBUG=$dbg timeout 600 python tools/microbench2.py > gpurun_out/microbench2_d$dbg.log 2>&1
grep -A40 "B=1 stage shapes" gpurun_out/microbench2_d$dbg.log | grep -v "d=3\|d=4\|act kernel\|C=...  *L=  *[0-9]*:" | head -24
done
bash tools/gpu_quick3.sh 2>&1 | grep -v "passed\|^\.\.\.\|kernel_shares\|{"
