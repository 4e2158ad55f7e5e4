#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 50 --warmup 5 --no-config5 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; cut -c1-200 gpurun_out/r2u_bench.json
timeout 600 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline --no-config5 --no-gpu-eager > gpurun_out/r2u_bench_sr48.json 2> gpurun_out/r2u_bench_sr48.err; cut -c1-200 gpurun_out/r2u_bench_sr48.json
BENCH_DUMP_KERNELS=gpurun_out/r2u_sr48_kernels.txt timeout 600 python bench.py --steps 3 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline --no-config5 --no-gpu-eager > /dev/null 2>&1; head -8 gpurun_out/r2u_sr48_kernels.txt
