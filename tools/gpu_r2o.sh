#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -q -x > gpurun_out/r2o_t.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2o_t.log
timeout 300 python tools/microbench_act.py 2>&1 | sed 's/tensor-core.*//' 
timeout 600 python bench.py --steps 20 --warmup 5 --no-config5 --no-cpu-baseline --no-gpu-eager --min-seconds 0.6 > gpurun_out/r2o_bench.log 2> gpurun_out/r2o_bench.err; echo "rc=$?"
python - <<PY
import json
l=[x for x in open("gpurun_out/r2o_bench.log") if x.startswith("{")]
j=json.loads(l[-1]); print("B=1:", round(j["value"],1), round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"],1), j["kernel_shares"])
PY
