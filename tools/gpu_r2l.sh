#!/bin/bash
mkdir -p gpurun_out
BENCH_DUMP_KERNELS=gpurun_out/r2l_synth_kernels.txt timeout 600 python bench.py --workload synth --steps 20 --warmup 5 --no-config5 --no-cpu-baseline --no-gpu-eager --min-seconds 0.5 > gpurun_out/r2l_bench_synth.log 2> gpurun_out/r2l_bench_synth.err; echo "synth rc=$?"
head -40 gpurun_out/r2l_synth_kernels.txt
python - <<PY
import json
l=[x for x in open("gpurun_out/r2l_bench_synth.log") if x.startswith("{")]
j=json.loads(l[-1]); print(round(j["ms_per_step"],4), j["launches_per_step"], j["kernel_shares"], {k:v for k,v in j["step_roofline"].items() if k in("kernel_sum_ms","gpu_busy_ms","kernels_in_graph","concurrency")})
PY
