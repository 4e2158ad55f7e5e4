#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of a workload (CUPTI through torch.profiler's chrome trace): start offset,
duration, stream, grid, block, registers, shared memory per kernel.  Debugging aid for the batch-1 critical path.

  python tools/timeline_b1.py [--workload vocoder] [--seconds 10] [--batch 1] [--out gpurun_out/timeline.txt]
"""
import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vocoder")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.txt"))
    ap.add_argument("--parallel-blocks", type=int, default=1)
    ap.add_argument("--pdl", type=int, default=1)
    a = ap.parse_args()
    import megatts2_hierspeechpp_b200 as hsv
    dev = torch.device("cuda:0")
    wl = bench.make_workload(a, 0)
    model = bench.build_model(wl, dev)
    for m in model.modules():
        if hasattr(m, "parallel_blocks"):
            m.parallel_blocks = bool(a.parallel_blocks)
    if not a.pdl:
        from megatts2_hierspeechpp_b200 import _lib
        _lib.load().hsv_set_pdl(0)
    runner = hsv.CudaGraphRunner(model)
    ins = [t.to(dev) for t in wl["host_inputs"]]
    for _ in range(5):
        runner(*ins)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            flush.fill_(1)
            torch.cuda.synchronize()
            runner(*ins)
            torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "t.json")
        prof.export_chrome_trace(p)
        tr = json.load(open(p))
    ks = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
    ks.sort(key=lambda e: e["ts"])
    # last replay = kernels after the last flush fill
    idx = [i for i, e in enumerate(ks) if "FillFunctor" in e["name"] and e["dur"] > 20]   # the 256 MB L2 flush
    ks = ks[idx[-1] + 1:] if idx else ks
    t0 = ks[0]["ts"]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        end = max(e["ts"] + e["dur"] for e in ks)
        f.write(f"# {wl['name']}: {len(ks)} kernels, span {end - t0:.1f} us\n")
        f.write("# start_us dur_us stream grid block regs smem name\n")
        for e in ks:
            g = e.get("args", {})
            f.write(f"{e['ts'] - t0:9.2f} {e['dur']:7.2f} s{g.get('stream', '?'):<3} grid={g.get('grid')} block={g.get('block')} "
                    f"regs={g.get('registers per thread')} smem={g.get('shared memory')} {e['name'][:70]}\n")
    print(open(a.out).read()[:300])


if __name__ == "__main__":
    main()
