#!/bin/bash
# bring-up of the swizzled-operand conv: descriptor variants, then the kernel tests under both layouts
mkdir -p gpurun_out
timeout 900 python tools/umma_diag.py > gpurun_out/umma_diag2.log 2>&1
cut -c1-330 gpurun_out/umma_diag2.log | grep -v '"max_err": [0-9.e-]*-0[5-9]' | tail -40
grep "====" gpurun_out/umma_diag2.log
HSV_LAYOUT=0 timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/t_layout0.log 2>&1; tail -3 gpurun_out/t_layout0.log
