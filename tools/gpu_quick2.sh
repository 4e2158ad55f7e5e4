#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48.log 2>&1
tail -2 gpurun_out/t_all.log
