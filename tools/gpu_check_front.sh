#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/chk_tests.log
cat gpurun_out/chk_tests.log
for w in synth tts; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --no-config5 > gpurun_out/chk_bench_$w.json 2> gpurun_out/chk_bench_$w.err
  cut -c1-200 gpurun_out/chk_bench_$w.json; tail -1 gpurun_out/chk_bench_$w.err
done
timeout 300 python tools/timeline_b1.py --workload synth --out gpurun_out/timeline_synth.txt > /dev/null 2>&1
