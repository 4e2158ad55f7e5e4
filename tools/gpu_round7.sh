#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/umma_diag.py > gpurun_out/umma_diag.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1
timeout 600 python tools/microbench2.py > gpurun_out/microbench2.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
HSV_UMMA_DEBUG=$((1<<24)) timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msub1.log 2>&1
HSV_UMMA_DEBUG=$((4<<24)) timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msub4.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48.log 2>&1
HSV_UMMA_DEBUG=$((4<<24)) timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48_msub4.log 2>&1
tail -3 gpurun_out/t_all.log
