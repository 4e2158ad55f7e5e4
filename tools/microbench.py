"""Steady-state per-launch timings of the two hot kernels (GPU box).  Back-to-back launches on one stream,
CUDA events around the batch, inputs cycled over > L2-size of distinct buffers when they are large."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402
from megatts2_hierspeechpp_b200 import _lib  # noqa: E402

dev = "cuda:0"
lib = _lib.load()


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n   # us per launch


print("== act1d (blk16 out): us/launch, GB/s (6 B/elem)  [variant 0 scalar | 1 packed]")
for (B, C, L) in [(1, 128, 1000), (1, 256, 2000), (1, 128, 10000), (1, 64, 40000), (1, 32, 80000), (1, 16, 160000),
                  (8, 16, 160000), (16, 32, 480000)]:
    nbuf = max(1, min(8, int(300e6 // (B * C * L * 4)) or 1))
    xs = [torch.randn(B, C, L, device=dev) for _ in range(nbuf)]
    a = torch.zeros(C, device=dev); b = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev)
    row = []
    for v in (0, 1):
        lib.hsv_set_act_variant(v)
        i = [0]

        def f():
            hsv.ops.act1d_blk16(xs[i[0] % nbuf], a, b, buf); i[0] += 1
        us = timeit(f)
        row.append((us, 6.0 * B * C * L / us / 1e3))
    print(f"  [{B},{C},{L}] " + " | ".join(f"{us:8.2f} us {gb:7.1f} GB/s" for us, gb in row))
    del xs
lib.hsv_set_act_variant(1)

print("== conv1d_umma (+bias +residual, fp32 out): us/launch, TFLOP/s, GB/s")
for (B, C, L, k, d) in [(1, 256, 2000, 11, 5), (1, 256, 2000, 3, 1), (1, 128, 10000, 11, 1), (1, 128, 10000, 3, 3),
                        (1, 64, 40000, 11, 5), (1, 64, 40000, 3, 1), (1, 32, 80000, 7, 3), (1, 16, 160000, 11, 1),
                        (8, 16, 160000, 7, 1), (16, 32, 480000, 7, 3), (16, 32, 480000, 11, 5)]:
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, k, device=dev) * 0.05
    bias = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev, slot=1)
    hsv.ops.pack_blk16(x, buf)
    nt = hsv.ops.pick_n_tile(C)
    wp = hsv.ops.pack_conv_weight(w, nt)
    out = torch.empty_like(x)
    us = timeit(lambda: hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=x, out=out), n=30)
    fl = 2.0 * B * C * C * k * L
    by = B * C * L * (2 + 4 + 4)
    print(f"  [{B},{C},{L}] k={k} d={d}: {us:8.2f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s")
    del x, out

print("== empty-ish launch overhead: add3_bcast on 1 element")
t = torch.zeros(1, 1, 1, device=dev)
print(f"  {timeit(lambda: hsv.ops.add3_bcast(t, None, None, out=t), n=200):.2f} us/launch")
