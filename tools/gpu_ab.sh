#!/bin/bash
mkdir -p gpurun_out
tail -2 gpurun_out/t_all.log 2>/dev/null
for rep in 1 2; do for dbg in 0 32; do
HSV_UMMA_DEBUG=$dbg timeout 300 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/ab_$dbg.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/ab_$dbg.log") if x.startswith("{")][-1]; j=json.loads(l)
print("debug=$dbg rep=$rep voc_b16", round(j["value"],1), round(j["ms_per_step"],3), j["kernel_shares"]["conv1d_umma"]["ms"], j["kernel_shares"]["act1d_blk16"]["ms"])
PY
done; done
