"""Sweep of the experimental operand paddings of the tcgen05 conv (set through HSV_UMMA_DEBUG)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import torch
    import megatts2_hierspeechpp_b200 as hsv
    sys.path.insert(0, os.path.join(ROOT, "tools"))
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
    res = []
    for (C, k) in ((128, 11), (64, 11), (256, 7)):
        x = torch.randn(1, C, 128, device=dev); w = torch.randn(C, C, k, device=dev) * 0.05
        buf = hsv.ops.blk16_buffer(1, C, 128, dev, slot=1); hsv.ops.pack_blk16(x, buf)
        nt = hsv.ops.pick_n_tile(C); wp = hsv.ops.pack_conv_weight(w, nt); out = torch.empty_like(x)
        ref = torch.nn.functional.conv1d(x.half().float(), w.half().float(), None, padding=(k - 1) // 2)
        y = hsv.ops.conv1d_umma(buf, wp, None, 128, C, C, k, 1, nt, out=out); torch.cuda.synchronize()
        err = (y - ref).abs().max().item()
        us = graph_time(lambda: hsv.ops.conv1d_umma(buf, wp, None, 128, C, C, k, 1, nt, out=out))
        res.append(f"C={C} k={k}: {us:6.2f} us err={err:.1e}")
    print(" | ".join(res), flush=True)
    sys.exit(0)
for apad in (0, 2, 4, 6):
    for bpad in (0, 2, 4, 6):
        flags = 256 | (apad << 16) | (bpad << 20)
        env = dict(os.environ, HSV_UMMA_DEBUG=str(flags))
        r = subprocess.run([sys.executable, __file__, "run"], capture_output=True, text=True, env=env, timeout=300)
        out = [l for l in r.stdout.splitlines() if l.startswith("C=")]
        print(f"apad={apad} bpad={bpad}: " + (out[0] if out else "FAIL " + r.stderr[-300:]), flush=True)
