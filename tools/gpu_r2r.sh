#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 --no-config5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-config5 --batch 16 > gpurun_out/r2q_bench_b16.json 2> gpurun_out/r2q_bench_b16.err
timeout 600 python bench.py --workload speechsr48 --batch 16 --steps 10 --warmup 3 --no-config5 > gpurun_out/r2q_bench_sr48.json 2> gpurun_out/r2q_bench_sr48.err
timeout 600 python bench.py --workload synth --steps 30 --warmup 5 --no-config5 > gpurun_out/r2q_bench_synth.json 2> gpurun_out/r2q_bench_synth.err
for f in r2q_bench r2q_bench_b16 r2q_bench_sr48 r2q_bench_synth; do cut -c1-300 gpurun_out/$f.json; tail -1 gpurun_out/$f.err; done
