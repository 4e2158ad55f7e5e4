#!/bin/bash
# round 2, call A: first contact of the tensor-core activation kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_act_mma.py -q -x > gpurun_out/r2a_t_actmma.log 2>&1; echo "actmma rc=$?"; tail -15 gpurun_out/r2a_t_actmma.log
timeout 300 python tools/microbench_act.py > gpurun_out/r2a_microbench_act.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/r2a_microbench_act.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2a_t_all.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/r2a_t_all.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?"
HSV_ACT_VARIANT=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_v1.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/r2a_bench_b16.log 2>&1
HSV_ACT_VARIANT=1 timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/r2a_bench_b16_v1.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/r2a_bench_sr48.log 2>&1
for f in r2a_bench r2a_bench_v1 r2a_bench_b16 r2a_bench_b16_v1 r2a_bench_sr48; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["value"],1), round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"],1))
    s=j["roofline_saturated"]; print("  sat act", round(s["act1d_kernel"]["frac"],3), "conv", round(s["conv_umma_kernel"]["frac"],3))
except Exception as e:
    print("$f", "ERR", e); print(open("gpurun_out/$f.log").read()[-1500:])
PY
done
