#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/r2g_t_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/r2g_t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config5 > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open("gpurun_out/r2g_bench.log") if x.startswith("{")]
j=json.loads(l[-1]); print(j["value"], j["ms_per_step"], j["e2e"]["value"], j["step_roofline"]["kernel_sum_ms"], j["clocks"])
PY
