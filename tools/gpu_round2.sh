#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "act1d" > gpurun_out/t_act.log 2>&1
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
HSV_ACT_VARIANT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scalar.log 2>&1
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_forward.csv python tools/profile_kernels.py forward > gpurun_out/p_forward.log 2>&1
tail -3 gpurun_out/t_act.log
