#!/bin/bash
mkdir -p gpurun_out
for pb in 1 0; do for B in 1 2 4; do
timeout 300 python bench.py --steps 20 --warmup 3 --batch $B --parallel-blocks $pb --no-cpu-baseline > gpurun_out/ab2.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/ab2.log") if x.startswith("{")][-1]; j=json.loads(l)
print("parallel_blocks=$pb B=$B", round(j["value"],1), "audio-s/s", round(j["ms_per_step"],3), "ms/step", round(j["ms_per_step"]/$B,3), "ms/utt")
PY
done; done
