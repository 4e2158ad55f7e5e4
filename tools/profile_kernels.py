"""Targets for ncu on the GPU box.

    python tools/profile_kernels.py forward            # 3 eager forwards of the B=1 x 10 s vocoder
    python tools/profile_kernels.py sr48 B             # 2 eager forwards of SpeechSR48, B x 10 s
    python tools/profile_kernels.py act  B C L [mode]  # isolated fused activation launches
    python tools/profile_kernels.py umma B C L k d     # isolated tcgen05 conv launches
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402
from megatts2_hierspeechpp_b200 import synthetic as synth  # noqa: E402

dev = "cuda:0"
what = sys.argv[1]
if what == "forward":
    m = hsv.Vocoder()
    m.load_state_dict(synth.vocoder_sd(1234), strict=True)
    m.to(dev).eval()
    z, g = synth.vocoder_inputs(1, 500, seed=1111)
    z, g = z.to(dev), g.to(dev)
    with torch.no_grad():
        for _ in range(3):
            m(z, g)
            torch.cuda.synchronize()
elif what == "sr48":
    import numpy as np
    B = int(sys.argv[2])
    sd = {k: torch.from_numpy(v.copy()) for k, v in np.load(os.path.join(ROOT, "tests/golden/speechsr48_state.npz")).items()}
    m = hsv.SpeechSR48(128, 40, **hsv.SR_CFG)
    m.load_state_dict(sd, strict=True)
    m.to(dev).eval()
    x = synth.speechsr_input(B, 160000).to(dev)
    with torch.no_grad():
        for _ in range(2):
            m(x)
            torch.cuda.synchronize()
elif what == "act":
    B, C, L = map(int, sys.argv[2:5])
    mode = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    x = torch.randn(B, C, L, device=dev)
    a = torch.zeros(C, device=dev); b = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev)
    out = torch.empty_like(x)
    for _ in range(5):
        if mode == 1:
            hsv.ops.act1d_blk16(x, a, b, buf)
        else:
            hsv.ops.act1d(x, a, b, out=out)
    torch.cuda.synchronize()
elif what == "umma":
    B, C, L, k, d = map(int, sys.argv[2:7])
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, k, device=dev) * 0.05
    bias = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev)
    hsv.ops.pack_blk16(x, buf)
    nt = hsv.ops.pick_n_tile(C, B * ((L + 127) // 128))
    wp = hsv.ops.pack_conv_weight(w, nt)
    out = torch.empty_like(x)
    for _ in range(5):
        hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=x, out=out)
    torch.cuda.synchronize()
