#!/bin/bash
mkdir -p gpurun_out
for v in 1 8449 4353; do
  HSV_ACT_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_actv$v.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_actv$v.log") if x.startswith("{")][-1]; j=json.loads(l)
print("act variant $v:", round(j["value"],1), round(j["ms_per_step"],4))
PY
done
HSV_ACT_VARIANT=8449 timeout 600 python tools/microbench2.py 2>&1 | grep -A8 "act kernel"
