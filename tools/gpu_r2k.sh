#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload synth --steps 20 --warmup 5 --no-config5 > gpurun_out/r2k_bench_synth.log 2> gpurun_out/r2k_bench_synth.err; echo "synth rc=$?"; tail -c 500 gpurun_out/r2k_bench_synth.err
timeout 600 python bench.py --workload chain24 --steps 20 --warmup 5 --no-config5 --no-cpu-baseline > gpurun_out/r2k_bench_chain.log 2> gpurun_out/r2k_bench_chain.err; echo "chain rc=$?"; tail -c 300 gpurun_out/r2k_bench_chain.err
for f in r2k_bench_synth r2k_bench_chain; do python - <<PY
import json
l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1])
    print("$f", round(j["value"],1), "ms", round(j["ms_per_step"],4), "launches", j["launches_per_step"], "e2e", round(j["e2e"]["value"],1), "pcm", round(j["e2e_pcm16"]["value"],1))
    print("  eager gpu", j.get("gpu_eager_baseline"), "cpu", j.get("cpu_baseline"))
    print("  shares", j["kernel_shares"]); print("  step", {k:v for k,v in j["step_roofline"].items() if k!="note"})
PY
done
