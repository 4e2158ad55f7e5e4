#!/bin/bash
mkdir -p gpurun_out
HSV_ACT_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:act1d_mma -s 2 -c 1 -f -o gpurun_out/r2_prof_actmma_sat python tools/profile_kernels.py act 16 32 480000 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
