#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 4 8 12; do echo "=== debug $dbg"; TRACE_SHORT=1 HSV_UMMA_DEBUG=$dbg timeout 200 python tools/umma_trace.py 2>&1 | grep -v "w_full\|a_full\|setup\|entry  \|producer"; done | tee gpurun_out/umma_trace_dbg.log
