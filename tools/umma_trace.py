"""Phase timing of ONE conv CTA (clock64 stamps written by CTA (0,0,0)): where a launch's latency goes.

    python tools/umma_trace.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402
from megatts2_hierspeechpp_b200 import _lib  # noqa: E402

dev = "cuda:0"
NAMES = ["entry", "setup done", "producer: pdl_wait done", "mma: w_full[0]", "mma: a_full[0]", "mma: all issued",
         "epi: pdl_wait done", "epi: acc_full", "epi: stores issued", "exit sync", "producer: before pdl_wait"]


def trace(B, C, L, k, d, residual=True, nt=None):
    lib = _lib.load()
    buf_t = torch.zeros(32, dtype=torch.int64, device=dev)
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, k, device=dev) * 0.05
    bias = torch.zeros(C, device=dev)
    hsv.ops.clear_workspace()
    buf = hsv.ops.blk16_buffer(B, C, L, dev, slot=1)
    hsv.ops.pack_blk16(x, buf)
    nt = nt or hsv.ops.pick_n_tile(C, B * ((L + 127) // 128))
    wp = hsv.ops.pack_conv_weight(w, nt)
    out = torch.empty_like(x)
    run = lambda: hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=x if residual else None, out=out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    lib.hsv_set_umma_trace(ctypes.c_void_p(buf_t.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.hsv_set_umma_trace(None)
    t = buf_t.cpu().tolist()
    t0 = t[0]
    print(f"C={C} L={L} B={B} k={k} d={d} n_tile={nt} res={int(residual)}  (SM cycles since entry; ~1.9 cycles/ns)")
    t1 = t[16]
    for i, n in enumerate(NAMES):
        print(f"   {n:28s} {t[i] - t0:8d}     mid-grid CTA: {t[16 + i] - t1 if t1 else 0:8d}")


if __name__ == "__main__":
    if os.environ.get("TRACE_SHORT"):
        trace(1, 128, 10000, 11, 5)
        trace(16, 128, 10000, 11, 5)
        trace(16, 256, 2000, 11, 5)
        trace(16, 64, 40000, 11, 5)
        trace(16, 32, 80000, 7, 3)
        sys.exit(0)
    trace(1, 256, 128, 1, 1)
    trace(1, 256, 128, 11, 1)
    trace(1, 256, 128, 1, 1, residual=False)
    trace(1, 128, 128, 1, 1)
    trace(1, 128, 128, 11, 1)
    trace(1, 64, 128, 1, 1)
    trace(1, 256, 2000, 11, 5)
    trace(1, 128, 10000, 11, 5)
    trace(1, 64, 40000, 11, 5)
    trace(1, 32, 80000, 11, 5)
    trace(16, 128, 10000, 11, 5)
