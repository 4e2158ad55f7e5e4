#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 900 python tools/microbench4.py > gpurun_out/microbench4.log 2>&1; cat gpurun_out/microbench4.log
