"""numpy model of the tensor-core activation kernel (csrc/act1d_mma.cu), step for step: same tiles, rows, K-blocks,
output windows, fp16 hi/lo splits and Toeplitz tables (tools/gen_act_tables.py).  Test infrastructure only: it
pins the table generator and the kernel's index arithmetic on CPU; the GPU tests compare the real kernel with the
oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_act_tables as T  # noqa: E402

RUNS, RT = 16, 32
WIN, VALID = RUNS * RT, (RUNS - 2) * RT


def _f16(a):
    return a.astype(np.float16).astype(np.float32)


def split_trunc(x):
    """hi = x with the low 13 mantissa bits cleared (exact in fp16 for |x| in the fp16 range), lo = fp16(x - hi)."""
    hi = (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = _f16(x - hi)
    return _f16(hi), lo


def act_tile(xrow, alpha, beta, L, tw, sc=1.0):
    """One channel row, one tile window [tw, tw+WIN): returns out[WIN] (runs 0 and 15 are garbage by design)."""
    tabs = {n: m.astype(np.float32) for n, m in T.tables()}
    t = np.clip(np.arange(tw, tw + WIN), 0, L - 1)
    x = (xrow[t] * np.float32(sc)).astype(np.float32)
    hi, lo = split_trunc(x)
    # XA rows: [run][64] interleaved (hi, lo); rows -1 and 16 are zero
    xa = np.zeros((RUNS + 2, 2 * RT), dtype=np.float32)
    xa[1:-1, 0::2] = hi.reshape(RUNS, RT)
    xa[1:-1, 1::2] = lo.reshape(RUNS, RT)
    # ---- up FIR: Y cols [-32, 96) stored at +32 ----
    Y = np.full((RUNS, 128), np.nan, dtype=np.float32)
    blocks = [(-1, 3), (0, 0), (0, 1), (0, 2), (0, 3), (1, 0)]   # (row shift in runs, K-step) for j = -1..4
    order = [1, 4, 0, 2, 3, 5]                                       # j = 0 and j = 3 overwrite first
    for n_done, bi in enumerate(order):
        shift, ks = blocks[bi]
        j = bi - 1
        a = xa[1 + shift:1 + shift + RUNS, 16 * ks:16 * ks + 16]      # [RUNS,16]
        contrib = a @ tabs["up_hi"].T + a @ tabs["up_lo"].T           # [RUNS,48]
        c0 = 16 * j - 16 + 32
        if n_done < 2:
            Y[:, c0:c0 + 48] = contrib
        else:
            Y[:, c0:c0 + 48] += contrib
    y = Y[:, 32:96]                                                   # 64 2x samples per run: idx = n - (2*t0 - 1)
    a_ = np.exp(np.float32(alpha)).astype(np.float32)
    ib = np.float32(1.0) / (np.exp(np.float32(beta)).astype(np.float32) + np.float32(1e-9))
    s = np.sin((y * a_).astype(np.float32)).astype(np.float32)
    z = (y + (ib * s) * s).astype(np.float32)
    # 2x-grid replicate clamp of the ACTIVATED signal
    n_idx = 2 * (tw + RT * np.arange(RUNS))[:, None] - 1 + np.arange(64)[None, :]
    if tw < 0 or tw + WIN > L:
        zL, zR = _edge_z(xrow, alpha, beta, L, sc)
        z = np.where(n_idx < 0, zL, np.where(n_idx > 2 * L - 1, zR, z)).astype(np.float32)
    zq = np.zeros((RUNS + 2, 64), dtype=np.float32)
    zq[1:-1] = _f16(z)
    # ---- down FIR: O cols [-16, 48) stored at +16 ----
    O = np.full((RUNS, 64), np.nan, dtype=np.float32)
    order = [0, 4, 1, 2, 3, 5]                                       # j = -1 and j = 3 overwrite first
    for n_done, bi in enumerate(order):
        shift, ks = blocks[bi]
        j = bi - 1
        a = zq[1 + shift:1 + shift + RUNS, 16 * ks:16 * ks + 16]
        even = (j % 2 == 0)
        contrib = a @ (tabs["dn_even"] if even else tabs["dn_odd"]).T
        w0 = (8 * j - 16) if even else (8 * j - 8)
        c0 = w0 + 16
        if n_done < 2:
            O[:, c0:c0 + 32] = contrib
        else:
            O[:, c0:c0 + 32] += contrib
    return O[:, 16:48].reshape(-1)


def _edge_z(xrow, alpha, beta, L, sc):
    f = T.TAPS.astype(np.float32)   # the edge samples come from the fp32 scalar path of the kernel

    def up(m, q):   # y[2m-1] (q=0) / y[2m] (q=1)
        acc = np.float32(0)
        for i in range(6):
            acc += np.float32(2) * f[2 * i + q] * np.float32(xrow[min(max(m + 2 - i, 0), L - 1)] * np.float32(sc))
        return np.float32(acc)

    a_ = np.float32(np.exp(np.float32(alpha)))
    ib = np.float32(1.0) / (np.float32(np.exp(np.float32(beta))) + np.float32(1e-9))

    def sn(y):
        s = np.float32(np.sin(np.float32(y * a_)))
        return np.float32(y + ib * s * s)

    return sn(up(0, 1)), sn(up(L, 0))     # z[0], z[2L-1]


def activation1d(x, alpha, beta, sc=1.0):
    """x [B,C,L] fp32 -> fp32 [B,C,L] following the kernel's tiling (values before the final fp16 rounding)."""
    B, C, L = x.shape
    out = np.zeros_like(x, dtype=np.float32)
    ntiles = (L + VALID - 1) // VALID
    for b in range(B):
        for c in range(C):
            for k in range(ntiles):
                tw = -RT + VALID * k
                o = act_tile(x[b, c], alpha[c], beta[c], L, tw, sc)
                lo, hi = tw + RT, min(tw + RT + VALID, L)
                out[b, c, lo:hi] = o[RT:RT + (hi - lo)]
    return out
