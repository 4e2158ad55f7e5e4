#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
BENCH_DUMP_LAUNCHES=gpurun_out/bench_launches.txt timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sw.log 2>&1
sort -k2 -n -r gpurun_out/bench_launches.txt | head -12
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:act1d -s 2 -c 1 -o gpurun_out/prof2_act_sat -f python tools/profile_kernels.py act 16 32 480000 1 > gpurun_out/p2_act.log 2>&1
$NCU --set full --import-source on -k regex:act1d -s 2 -c 1 -o gpurun_out/prof2_act_b1 -f python tools/profile_kernels.py act 1 16 160000 1 > gpurun_out/p2_act1.log 2>&1
ls -la gpurun_out/prof2_act*
