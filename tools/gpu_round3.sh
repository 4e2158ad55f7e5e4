#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_forward.csv python tools/profile_kernels.py forward > gpurun_out/p_forward.log 2>&1
tail -3 gpurun_out/t_all.log
