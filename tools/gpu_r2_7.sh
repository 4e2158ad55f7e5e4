#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof2_umma_c128_b1 -f python tools/profile_kernels.py umma 1 128 10000 11 5 > gpurun_out/p2_umma1.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof2_umma_c128_b16 -f python tools/profile_kernels.py umma 16 128 10000 11 5 > gpurun_out/p2_umma2.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof2_umma_c32_b16 -f python tools/profile_kernels.py umma 16 32 80000 7 3 > gpurun_out/p2_umma3.log 2>&1
ls -la gpurun_out/prof2*
