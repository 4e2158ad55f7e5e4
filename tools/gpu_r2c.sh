#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_act_mma.py -q -x > gpurun_out/r2p_t_actmma.log 2>&1; echo "actmma rc=$?"; tail -5 gpurun_out/r2p_t_actmma.log
timeout 300 python tools/microbench_act.py > gpurun_out/r2p_microbench_act.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/r2p_microbench_act.log
HSV_ACT_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:act1d_mma -s 2 -c 1 -f -o gpurun_out/r2_prof_actmma_sat4 python tools/profile_kernels.py act 16 32 480000 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
