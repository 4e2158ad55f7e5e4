#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config5 --no-gpu-eager --min-seconds 0.6 > gpurun_out/r2i_$name.log 2>/dev/null
  python - <<PY
import json
l=[x for x in open("gpurun_out/r2i_$name.log") if x.startswith("{")]
j=json.loads(l[-1]); print("$name", round(j["value"],1), round(j["ms_per_step"],4), j["launches_per_step"], round(j["step_roofline"]["kernel_sum_ms"],3))
PY
}
run base HSV_FUSE_MAX_C=0
run f64_all HSV_FUSE_MAX_C=64
run f64_300k HSV_FUSE_MAX_C=64 HSV_FUSE_MAX_ELEMS=300000
run f64_1m HSV_FUSE_MAX_C=64 HSV_FUSE_MAX_ELEMS=1400000
run f64_3m HSV_FUSE_MAX_C=64 HSV_FUSE_MAX_ELEMS=3000000
run f32_3m HSV_FUSE_MAX_C=32 HSV_FUSE_MAX_ELEMS=3000000
