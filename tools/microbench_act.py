"""Fused activation: CUDA-core kernel (act1d.cu) vs tensor-core kernel (act1d_mma.cu), graph-timed, saturated and
batch-1 shapes.  fp32 [B,C,L] in, fp16 blk16 out: 6 algorithmic bytes per element."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402
from megatts2_hierspeechpp_b200 import _lib  # noqa: E402
dev = "cuda:0"


def graph_time(fn, n=10, reps=3):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


lib = _lib.load()
for (B, C, L) in ((16, 32, 480000), (16, 64, 40000), (16, 256, 2000), (16, 16, 160000), (1, 16, 160000), (1, 32, 80000),
                  (1, 64, 40000), (1, 128, 10000), (1, 256, 2000), (1, 128, 1000)):
    x = torch.randn(B, C, L, device=dev)
    a = torch.zeros(C, device=dev)
    hsv.ops.clear_workspace()
    buf = hsv.ops.blk16_buffer(B, C, L, dev)
    line = f"[{B},{C},{L}]"
    for name, v in (("cuda-core", 1), ("tensor-core", 2)):
        lib.hsv_set_act_variant(v)
        us = graph_time(lambda: hsv.ops.act1d_blk16(x, a, a, buf))
        line += f"  {name}: {us:8.2f} us ({6.0 * B * C * L / us / 1e3:6.0f} GB/s)"
    print(line, flush=True)
lib.hsv_set_act_variant(0)
