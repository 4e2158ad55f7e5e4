"""Tile-policy sweep of the tcgen05 conv at batch scale: (n_tile, msub) per layer shape, graph-timed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402

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


def case(B, C, L, k, d, nt, msub):
    hsv.ops.clear_workspace()
    hsv.ops.set_umma_debug(msub << 24)
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, k, device=dev) * 0.05
    bias = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev, slot=1)
    hsv.ops.pack_blk16(x, buf)
    wp = hsv.ops.pack_conv_weight(w, nt)
    out = torch.empty_like(x)
    try:
        us = graph_time(lambda: hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=x, out=out))
    except Exception as e:  # e.g. shared memory exceeded
        print(f"  B={B} C={C:3d} L={L:6d} k={k:2d} n_tile={nt:3d} msub={msub}: {str(e)[:80]}")
        return
    tf = 2.0 * B * C * C * k * L / us / 1e6
    gb = (10.0 * B * C * L) / us / 1e3
    print(f"  B={B} C={C:3d} L={L:6d} k={k:2d} n_tile={nt:3d} msub={msub}: {us:8.2f} us  {tf:7.1f} TFLOP/s  {gb:7.1f} GB/s", flush=True)


SHAPES = ((256, 2000, (256, 128, 64)), (128, 10000, (128, 64)), (64, 40000, (64,)), (32, 80000, (32,)))
if len(sys.argv) > 1 and sys.argv[1] == "small":
    SHAPES = ((16, 160000, (16,)), (32, 80000, (32,)))
for B in (16, 1):
    print(f"== B={B}")
    for (C, L, nts) in SHAPES:
        for k, d in ((3, 1), (7, 1), (11, 5)):
            for nt in nts:
                for msub in (1, 2, 4):
                    if B == 1 and msub == 4 and C > 32:
                        continue
                    case(B, C, L, k, d, nt, msub)
hsv.ops.set_umma_debug(0)
