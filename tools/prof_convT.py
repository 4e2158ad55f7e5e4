#!/usr/bin/env python
"""Profiling target: one polyphase ConvTranspose1d launch of the tcgen05 kernel.  python tools/prof_convT.py B Cin Cout k u L"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megatts2_hierspeechpp_b200 as hsv
ops = hsv.ops
B, cin, cout, k, u, L = (int(a) for a in sys.argv[1:7])
dev = "cuda:0"
x = torch.randn(B, cin, L, device=dev)
w = torch.randn(cin, cout, k, device=dev) * 0.05
bias = torch.randn(cout, device=dev)
buf = ops.blk16_buffer(B, cin, L, dev, 0)
ops.pack_blk16(x, buf)
rt = B * ((L + 127) // 128)
nt = ops.pick_n_tile(cout, rt * u)
wp = ops.pack_convT_weight(w, u, nt)
for _ in range(3):
    y = ops.conv_transpose1d_umma(buf, wp, bias, L, cin, cout, k, u, nt)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(gr, stream=side):
        for _ in range(20):
            ops.conv_transpose1d_umma(buf, wp, bias, L, cin, cout, k, u, nt)
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print(f"convT B={B} {cin}->{cout} k={k} u={u} L={L} n_tile={nt}: {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per launch (graph, L2 warm)")
