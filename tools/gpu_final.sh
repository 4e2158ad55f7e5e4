#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_voc_b16.log 2>&1
for f in bench bench_voc_b16; do python - <<PY
import json
l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
print("$f", round(j["value"],1), round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"],1), "clocks", j["clocks"], "cpu", j["cpu_baseline"] and j["cpu_baseline"]["value"])
PY
done
