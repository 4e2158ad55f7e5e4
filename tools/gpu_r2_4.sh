#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 300 python tools/umma_trace.py > gpurun_out/umma_trace.log 2>&1; grep -A6 "^C=" gpurun_out/umma_trace.log | grep -v "entry\|setup\|producer\|w_full" 
timeout 600 python tools/microbench2.py > gpurun_out/microbench2_sw.log 2>&1
grep -B100 "act kernel" gpurun_out/microbench2_sw.log | grep -v "d=3\|d=4"
