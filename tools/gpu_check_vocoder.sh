#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/chk_tests.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-config5 > gpurun_out/chk_bench.json 2> gpurun_out/chk_bench.err
timeout 200 python tools/timeline_b1.py --out gpurun_out/chk_timeline.txt > /dev/null 2>&1
cat gpurun_out/chk_tests.log; cut -c1-400 gpurun_out/chk_bench.json
