#!/bin/bash
# short refresh of the headline artifacts at the final HEAD (the full set: tools/gpu_artifacts_r2.sh)
mkdir -p gpurun_out
P=gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu -s > ${P}_t_all.log 2>&1; echo "tests rc=$?"; tail -2 ${P}_t_all.log; grep "\[parity\]" ${P}_t_all.log > ${P}_parity.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > ${P}_bench.log 2> ${P}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline --no-config5 > ${P}_bench_sr48_b16.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline --no-config5 > ${P}_bench_voc_b16.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --workload synth --no-config5 > ${P}_bench_synth.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --workload tts --no-config5 > ${P}_bench_tts.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --workload chain24 --no-config5 > ${P}_bench_chain24.log 2>&1
for f in bench bench_sr48_b16 bench_voc_b16 bench_synth bench_tts bench_chain24; do grep '^{' ${P}_$f.log | tail -1 | cut -c1-170; done
