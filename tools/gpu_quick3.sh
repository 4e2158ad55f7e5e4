#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sw.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48_sw.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_voc_b16_sw.log 2>&1
for f in bench_sw bench_sr48_sw bench_voc_b16_sw; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["value"],1), j["ms_per_step"], "e2e", round(j["e2e"]["value"],1)); print(json.dumps(j["kernel_shares"]))
    s=j["roofline_saturated"]; print("  sat act", round(s["act1d_kernel"]["frac"],3), "conv", round(s["conv_umma_kernel"]["frac"],3), " roofline", round(j["roofline"]["frac"],3), j["roofline"]["avg_launch_us"])
except Exception as e:
    print("$f", "ERR", e); print(open("gpurun_out/$f.log").read()[-1500:])
PY
done
