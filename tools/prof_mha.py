#!/usr/bin/env python
"""Profiling target: the attention kernel alone at the synthesizer's shape (B=1, 2 heads, D=96, T=500)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megatts2_hierspeechpp_b200 as hsv

ops = hsv.ops
B, H, D, T = 1, 2, 96, 500
C = H * D
qkv = torch.randn(B, 3 * C, T, device="cuda")
flat = qkv.view(-1)
for v in ([0, 1] if len(sys.argv) < 2 else [int(sys.argv[1])]):
    ops.set_mha_variant(v)
    for _ in range(3):
        out = ops.mha(flat, flat[C * T:], flat[2 * C * T:], B, H, D, T, T, 3 * C * T, 3 * C * T, 3 * C * T, D ** -0.5, False)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for _ in range(50):
                ops.mha(flat, flat[C * T:], flat[2 * C * T:], B, H, D, T, T, 3 * C * T, 3 * C * T, 3 * C * T, D ** -0.5, False)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"variant {v}: {e0.elapsed_time(e1) / 50 * 1e3:.2f} us per call (50 calls in one CUDA graph, L2 warm)")
