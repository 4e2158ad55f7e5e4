#!/usr/bin/env python
"""CUPTI kernel-time sums by kernel name over ONE eager single-stream forward of the vocoder (works in any checkout of the
repo: depends on the package only).  python tools/kernel_sums.py [batch]"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv
from megatts2_hierspeechpp_b200 import synthetic as synth
from torch.profiler import ProfilerActivity, profile
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda:0"
m = hsv.Vocoder(); m.load_state_dict(synth.vocoder_sd(1234), strict=True); m.to(dev).eval()
z, g = synth.vocoder_inputs(B, 500, seed=1111); z, g = z.to(dev), g.to(dev)
with torch.no_grad():
    for _ in range(3):
        m(z, g)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(z, g); torch.cuda.synchronize()
agg = collections.OrderedDict(); rows = []
for e in prof.events():
    if "cuda" in str(getattr(e, "device_type", "")).lower():
        n = __import__("re").sub(r"\(anonymous namespace\)::|void |<unnamed>::", "", e.name).split("(")[0][:60]
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
        rows.append((e.time_range.start, e.time_range.end - e.time_range.start, n))
tot = sum(v[1] for v in agg.values())
print(f"B={B}: total kernel time {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} kernels")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"{t/1e3:9.3f} ms  n={c:4d}  avg={t/c:8.1f} us  {n}")
rows.sort()
# the 12 longest individual launches
for st, du, n in sorted(rows, key=lambda r: -r[1])[:12]:
    print(f"   {du:8.1f} us  {n}")
