#!/bin/bash
# bench.py shake-out: N=1 default line (+config5), N=2 line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_reference.py -q -x -s > gpurun_out/r2f_t_ref.log 2>&1; echo "ref tests rc=$?"; grep -E "parity|passed|failed" gpurun_out/r2f_t_ref.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2f_bench.err
python - <<PY
import json
l=[x for x in open("gpurun_out/r2f_bench.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1])
    for k in ("value","ms_per_step","launches_per_step","clocks","step_roofline","config5","gpu_eager_baseline","cpu_baseline","kernel_shares"):
        print(k, json.dumps(j.get(k))[:900])
    print("roofline", json.dumps({k:v for k,v in j["roofline"].items() if k not in ("kernel","bound_note","timing")})[:900])
    print("e2e", j["e2e"])
PY
