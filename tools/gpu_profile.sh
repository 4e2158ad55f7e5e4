#!/bin/bash
# ncu session (1 GPU): launch list of the default bench command + full captures of the two hot kernels.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/p_bench.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_forward.csv python tools/profile_kernels.py forward > gpurun_out/p_forward.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_sr48.csv python tools/profile_kernels.py sr48 8 > gpurun_out/p_sr48.log 2>&1
$NCU --set full --import-source on -k regex:act1d -s 2 -c 1 -o gpurun_out/prof_act_b1 -f python tools/profile_kernels.py act 1 16 160000 1 > gpurun_out/p_act1.log 2>&1
$NCU --set full --import-source on -k regex:act1d -s 2 -c 1 -o gpurun_out/prof_act_sat -f python tools/profile_kernels.py act 16 32 480000 1 > gpurun_out/p_act2.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_umma_c256 -f python tools/profile_kernels.py umma 1 256 2000 11 5 > gpurun_out/p_umma1.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_umma_c32 -f python tools/profile_kernels.py umma 16 32 480000 7 3 > gpurun_out/p_umma2.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_umma_c128 -f python tools/profile_kernels.py umma 1 128 10000 11 1 > gpurun_out/p_umma3.log 2>&1
ls -la gpurun_out | tail -20
