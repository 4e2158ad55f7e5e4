#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_front.py -q -x -s > gpurun_out/r2m_t_front.log 2>&1; echo "front rc=$?"; grep -E "parity|passed|failed|Error|^E" gpurun_out/r2m_t_front.log | head -20
BENCH_DUMP_KERNELS=gpurun_out/r2m_synth_kernels.txt timeout 600 python bench.py --workload synth --steps 20 --warmup 5 --no-config5 --no-cpu-baseline --no-gpu-eager --min-seconds 0.5 > gpurun_out/r2m_bench_synth.log 2> gpurun_out/r2m_bench_synth.err; echo "synth rc=$?"
head -14 gpurun_out/r2m_synth_kernels.txt
python - <<PY
import json
l=[x for x in open("gpurun_out/r2m_bench_synth.log") if x.startswith("{")]
j=json.loads(l[-1]); print(round(j["ms_per_step"],4), j["launches_per_step"], {k:v for k,v in j["step_roofline"].items() if k in("kernel_sum_ms","gpu_busy_ms","kernels_in_graph","concurrency")})
PY
