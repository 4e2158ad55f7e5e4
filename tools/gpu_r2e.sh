#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/r2e_t_all.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/r2e_t_all.log; grep "\[parity\]" gpurun_out/r2e_t_all.log > gpurun_out/r2e_parity.log; wc -l gpurun_out/r2e_parity.log
