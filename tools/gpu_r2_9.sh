#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "fused or pcm16" > gpurun_out/t_fused.log 2>&1; tail -15 gpurun_out/t_fused.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
for fc in 0 32 64; do
  HSV_FUSE_MAX_C=$fc timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f$fc.log 2>&1
  HSV_FUSE_MAX_C=$fc timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48_f$fc.log 2>&1
  HSV_FUSE_MAX_C=$fc timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_voc_b16_f$fc.log 2>&1
  for f in bench_f$fc bench_sr48_f$fc bench_voc_b16_f$fc; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["value"],1), round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"],1), "launches", j["launches_per_step"])
except Exception as e:
    print("$f", "ERR", e); print(open("gpurun_out/$f.log").read()[-800:])
PY
  done
done
