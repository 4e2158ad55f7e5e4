#!/bin/bash
# round 2p: TTV tail (W2VDecoder + PitchPredictor) parity + tts chain bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ttv.py tests/test_gpu_front.py -x -q -m gpu -s 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2p_tests.log
timeout 600 python bench.py --workload tts --seconds 10 --steps 30 --warmup 5 --no-config5 > gpurun_out/r2p_bench_tts.json 2> gpurun_out/r2p_bench_tts.err
timeout 600 python bench.py --workload synth --seconds 10 --steps 30 --warmup 5 --no-config5 > gpurun_out/r2p_bench_synth.json 2> gpurun_out/r2p_bench_synth.err
tail -5 gpurun_out/r2p_tests.log; cat gpurun_out/r2p_bench_tts.json | cut -c1-600; tail -3 gpurun_out/r2p_bench_tts.err
