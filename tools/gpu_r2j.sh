#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_front.py -q -x -s > gpurun_out/r2j_t_front.log 2>&1; echo "front rc=$?"; grep -E "parity|passed|failed|Error|^E" gpurun_out/r2j_t_front.log | head -30
timeout 600 python -m pytest tests/test_gpu_reference.py -q -x -s > gpurun_out/r2j_t_ref.log 2>&1; echo "ref rc=$?"; grep -E "parity|passed|failed|^E" gpurun_out/r2j_t_ref.log | head
