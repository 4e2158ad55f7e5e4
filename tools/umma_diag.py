"""Bring-up probe for the tcgen05 conv (GPU box): structured operands expose descriptor/layout errors.

    python tools/umma_diag.py            # runs every probe in a subprocess (a trap cannot kill the rest)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def probe(name, C, k, d, L, mode):
    import torch
    import torch.nn.functional as F
    import megatts2_hierspeechpp_b200 as hsv

    dev = "cuda:0"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, C, L, generator=g).half().float()
    if mode == "identity":      # centre tap = I  -> out == x
        w = torch.zeros(C, C, k); w[range(C), range(C), (k - 1) // 2] = 1.0
    elif mode == "tap0":        # first tap = I   -> out[t] = x[t - h]
        w = torch.zeros(C, C, k); w[range(C), range(C), 0] = 1.0
    elif mode == "chan":        # out[co] = x[ci = (co+1) % C] at centre tap
        w = torch.zeros(C, C, k); w[range(C), [(c + 1) % C for c in range(C)], (k - 1) // 2] = 1.0
    else:
        w = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
    ref = F.conv1d(x.double(), w.double(), None, padding=(k - 1) // 2 * d, dilation=d)
    buf = hsv.ops.blk16_buffer(1, C, L, dev, slot=5)
    hsv.ops.pack_blk16(x.to(dev), buf)
    nt = hsv.ops.pick_n_tile(C)
    wp = hsv.ops.pack_conv_weight(w.to(dev), nt)
    y = hsv.ops.conv1d_umma(buf, wp, None, L, C, C, k, d, nt)
    torch.cuda.synchronize()
    y = y.cpu().double()
    err = (y - ref).abs()
    res = {"name": name, "C": C, "k": k, "d": d, "L": L, "mode": mode, "max_err": float(err.max()),
           "ref_max": float(ref.abs().max()), "frac_bad": float((err > 1e-3).float().mean())}
    if err.max() > 1e-3:
        res["y_head"] = y[0, :4, :6].tolist()
        res["ref_head"] = ref[0, :4, :6].tolist()
        bad = (err > 1e-3).nonzero()
        res["first_bad"] = bad[:8].tolist()
        res["bad_rows_mod128"] = sorted(set((bad[:, 2] % 128).tolist()))[:16]
        res["bad_ch"] = sorted(set(bad[:, 1].tolist()))[:16]
    print("PROBE " + json.dumps(res), flush=True)


CASES = [("gemm16_id", 16, 1, 1, 128, "identity"), ("gemm16_chan", 16, 1, 1, 128, "chan"),
         ("gemm16_rand", 16, 1, 1, 256, "rand"), ("gemm32_chan", 32, 1, 1, 128, "chan"),
         ("gemm64_chan", 64, 1, 1, 128, "chan"), ("gemm64_rand", 64, 1, 1, 256, "rand"),
         ("gemm128_rand", 128, 1, 1, 256, "rand"),
         ("c64_k3_d8", 64, 3, 8, 256, "rand"),       # row shifts that are multiples of 8 (same swizzle phase)
         ("c64_k3_tap0", 64, 3, 1, 256, "tap0"),     # shift by one row
         ("c32_k3_tap0", 32, 3, 1, 256, "tap0"), ("c16_k3_tap0", 16, 3, 1, 256, "tap0"),
         ("k3_d3_tap0", 16, 3, 3, 256, "tap0"), ("c32_rand", 32, 7, 3, 300, "rand"),
         ("c64_rand", 64, 11, 5, 300, "rand"), ("c128_rand", 128, 7, 1, 300, "rand"),
         ("c256_rand", 256, 11, 5, 300, "rand")]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        name, C, k, d, L, mode = sys.argv[1], *map(int, sys.argv[2:6]), sys.argv[6]
        probe(name, C, k, d, L, mode)
        sys.exit(0)
    for dbg in (os.environ.get("HSV_DIAG_MODES", "0,1,2").split(",")):
        print(f"==== HSV_UMMA_DEBUG={dbg}", flush=True)
        n_ok = 0
        for c in CASES:
            env = dict(os.environ, HSV_UMMA_DEBUG=dbg)
            try:
                r = subprocess.run([sys.executable, __file__, *map(str, c)], capture_output=True, text=True, env=env,
                                   timeout=120)
                out = [l for l in r.stdout.splitlines() if l.startswith("PROBE")]
                print(out[0] if out else f"FAIL {c[0]} rc={r.returncode} {r.stdout[-200:]} {r.stderr[-300:]}", flush=True)
                if out and json.loads(out[0][6:])["max_err"] < 1e-3:
                    n_ok += 1
            except subprocess.TimeoutExpired:
                print(f"TIMEOUT {c[0]}", flush=True)
        print(f"==== variant {dbg}: {n_ok}/{len(CASES)} ok", flush=True)
        if n_ok == len(CASES):
            break
