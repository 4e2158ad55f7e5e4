"""Per-launch GPU time of the tcgen05 conv for controlled shapes: 20 launches captured in a CUDA graph
(removes the Python launch overhead), weights L2-hot."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402

dev = "cuda:0"


def graph_time(fn, n=20, reps=5):
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


def conv_case(B, C, L, k, d, residual=True):
    hsv.ops.clear_workspace()
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, k, device=dev) * 0.05
    bias = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev, slot=1)
    hsv.ops.pack_blk16(x, buf)
    nt = hsv.ops.pick_n_tile(C, B * ((L + 127) // 128)) if L > 128 else hsv.ops.pick_n_tile(C)
    wp = hsv.ops.pack_conv_weight(w, nt)
    out = torch.empty_like(x)
    us = graph_time(lambda: hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=x if residual else None, out=out))
    ks = k * C // 16
    tiles = (L + 127) // 128
    print(f"  C={C:3d} L={L:6d} k={k:2d} d={d} res={int(residual)}: {us:7.2f} us  ksteps/CTA={ks:4d}  tiles={tiles:4d}x{C // nt}  "
          f"{2.0 * B * C * C * k * L / us / 1e6:7.1f} TFLOP/s", flush=True)


print("== single tile (one CTA per n-tile): latency vs K depth and tap alignment")
for C in (256, 128, 64):
    for (k, d) in ((1, 1), (3, 1), (7, 1), (7, 3), (7, 4), (11, 1)):
        conv_case(1, C, 128, k, d)
print("== no residual")
conv_case(1, 256, 128, 7, 1, residual=False)
conv_case(1, 128, 128, 7, 1, residual=False)
print("== B=1 stage shapes")
for (C, L) in ((256, 2000), (128, 10000), (64, 40000), (32, 80000), (16, 160000)):
    for (k, d) in ((3, 1), (11, 5)):
        conv_case(1, C, L, k, d)
print("== B=16 stage shapes (throughput regime)")
for (C, L) in ((256, 2000), (128, 10000), (64, 40000), (32, 80000)):
    for (k, d) in ((3, 1), (7, 3), (11, 5)):
        conv_case(16, C, L, k, d)
print("== act kernel (blk16 out), graph-timed")
for (C, L) in ((128, 1000), (256, 2000), (128, 10000), (64, 40000), (32, 80000), (16, 160000)):
    x = torch.randn(1, C, L, device=dev)
    a = torch.zeros(C, device=dev); b = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(1, C, L, dev)
    us = graph_time(lambda: hsv.ops.act1d_blk16(x, a, b, buf))
    print(f"  C={C:3d} L={L:6d}: {us:7.2f} us  {6.0 * C * L / us / 1e3:7.1f} GB/s", flush=True)
t = torch.zeros(1, 1, 1, device=dev)
print(f"== trivial kernel in graph: {graph_time(lambda: hsv.ops.add3_bcast(t, None, None, out=t)):.2f} us")
