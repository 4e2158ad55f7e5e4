#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 600 python tools/microbench2.py > gpurun_out/microbench2_sw.log 2>&1
HSV_UMMA_DEBUG=$((2<<24)) timeout 600 python tools/microbench2.py > gpurun_out/microbench2_sw_msub2.log 2>&1
HSV_UMMA_DEBUG=$((1<<24)) timeout 600 python tools/microbench2.py > gpurun_out/microbench2_sw_msub1.log 2>&1
grep -B100 "act kernel" gpurun_out/microbench2_sw.log
echo "---- msub2"; grep -A13 "B=16" gpurun_out/microbench2_sw_msub2.log
echo "---- msub1"; grep -A13 "B=16" gpurun_out/microbench2_sw_msub1.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sw.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48_sw.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_voc_b16_sw.log 2>&1
for f in bench_sw bench_sr48_sw bench_voc_b16_sw; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["value"],1), j["ms_per_step"], "e2e", round(j["e2e"]["value"],1)); print(json.dumps(j["kernel_shares"]))
except Exception as e:
    print("$f", "ERR", e); print(open("gpurun_out/$f.log").read()[-1500:])
PY
done
