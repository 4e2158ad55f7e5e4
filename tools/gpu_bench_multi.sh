#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi_bench_n$N.log 2> gpurun_out/multi_bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 600 gpurun_out/multi_bench_n$N.err
python - <<PY
import json
l=[x for x in open("gpurun_out/multi_bench_n$N.log") if x.startswith("{")]
j=json.loads(l[-1]); print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "clocks", j["clocks"])
print(json.dumps(j["config5"], indent=1)[:2500])
PY
timeout 300 python -m pytest tests/test_gpu_modules.py -q -k two_gpu -s 2>&1 | tail -3
