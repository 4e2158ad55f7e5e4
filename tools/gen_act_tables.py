#!/usr/bin/env python
"""Generate ``csrc/act_toeplitz_tables.h``: the constant B operands of the tensor-core FIRs of the fused
activation (``csrc/act1d_mma.cu``).

Activation1d's two 12-tap FIRs (alias_free_torch/resample.py:25-32 up, filter.py:86-94 down) are evaluated as
banded-Toeplitz MMAs.  One MMA row is one (channel, run) pair holding RT = 32 consecutive time steps; every
K-step of 16 fp16 values touches a shift-invariant window of outputs, so ONE small [N x 16] matrix per FIR
(and per window alignment) serves every K-step:

  up   A row = 32 steps x (hi, lo) fp16 split of x -> K = 64; K-block j (8 steps) feeds the 2x-rate output
       columns [16j-16, 16j+32) (N = 48); coefficient of x[8j+tau] in column n' is
       2*f[2*i + q], i = (n'>>1) - 6 - tau, q = n' & 1 (q = 0: odd sample y[2m-1], q = 1: even sample y[2m]).
  down A row = 64 fp16 2x-rate samples z[2*t0-1 .. 2*t0+62]; K-block j (16 samples) feeds the outputs
       [w0, w0+32) with w0 = 8j-16 (j even) or 8j-8 (j odd); coefficient of sample kappa in output n' is
       f[36 + kappa - 2n'] (even) / f[20 + kappa - 2n'] (odd).

Tap precision.  The up-sampling FIR feeds the non-linearity (errors in y are amplified by up to 1 + e^alpha / e^beta,
~200 in the bundled checkpoints), so its taps are split into fp16 hi + fp16 lo (two MMAs per K-block: (xh + xl) * Uh +
xh * Ul, ~22 significant bits).  The low-pass FIR after the non-linearity is linear and its result is rounded to fp16
anyway: its taps are rounded to fp16 ONCE, such that both polyphase sums stay exactly 0.5 (unit DC gain as in fp32;
largest deviation from the fp32 taps 1.2 fp16 ulp; one MMA per K-block).  Each matrix is stored in tcgen05's K-major SWIZZLE_32B canonical layout (32-byte rows; byte o of row-linear order
stored at o ^ (((o >> 7) & 1) << 4)), exactly the form of a 16-channel weight block of the conv kernel.

Run:  python tools/gen_act_tables.py            (rewrites the header)
      python tools/gen_act_tables.py --check    (exit 1 if the committed header is stale)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "megatts2_hierspeechpp_b200", "csrc", "act_toeplitz_tables.h")

TAPS = np.array([0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971,
                 0.4432097971, 0.1285725832, -0.0576573834, -0.0255434588, 0.0093894657, 0.0020289647],
                dtype=np.float32)

UP_N, DN_N, KB = 48, 32, 16


def fp16_taps() -> np.ndarray:
    """The 12 taps in fp16 with both polyphase sums exactly 0.5: start from round-to-nearest and move single taps by
    one ulp (never further than 1.5 ulp from the fp32 value) while that reduces the DC error.  Deterministic."""
    g = TAPS.astype(np.float16)
    for ph in (0, 1):
        idx = np.arange(ph, 12, 2)
        for _ in range(200):
            err = 0.5 - g[idx].astype(np.float64).sum()
            best = None
            for i in idx:
                up = np.nextafter(g[i], np.float16(np.inf))
                dn = np.nextafter(g[i], np.float16(-np.inf))
                for cand in (up, dn):
                    ne = abs(err - (float(cand) - float(g[i])))
                    if abs(float(cand) - float(TAPS[i])) <= 1.5 * abs(float(up) - float(g[i])) and \
                            (best is None or ne < best[0]):
                        best = (ne, i, cand)
            if best is None or best[0] >= abs(err):
                break
            g[best[1]] = best[2]
    assert np.array_equal(g, g[::-1]), "taps must stay symmetric"
    return g.astype(np.float32)


TAPS16 = fp16_taps()


def up_matrix() -> np.ndarray:
    """[48, 16] fp32: column n' x K index kk (kk = 2*tau + part; the same coefficient for the hi and lo part)."""
    m = np.zeros((UP_N, KB), dtype=np.float32)
    for n in range(UP_N):
        for kk in range(KB):
            tau, q = kk >> 1, n & 1
            i = (n >> 1) - 6 - tau
            if 0 <= i <= 5:
                m[n, kk] = np.float32(2.0) * TAPS[2 * i + q]     # x2 is exact
    return m


def down_matrix(even: bool) -> np.ndarray:
    m = np.zeros((DN_N, KB), dtype=np.float32)
    base = 36 if even else 20
    for n in range(DN_N):
        for kappa in range(KB):
            jj = base + kappa - 2 * n
            if 0 <= jj <= 11:
                m[n, kappa] = TAPS16[jj]
    return m


def swizzle32(m16: np.ndarray) -> np.ndarray:
    """[N,16] fp16 -> uint16 array in K-major SWIZZLE_32B physical order."""
    n = m16.shape[0]
    out = np.zeros(n * KB, dtype=np.uint16)
    bits = m16.view(np.uint16)
    for r in range(n):
        for k in range(KB):
            lin = r * 32 + (k >> 3) * 16
            phys = (lin ^ (((lin >> 7) & 1) << 4)) + 2 * (k & 7)
            out[phys >> 1] = bits[r, k]
    return out


def split(m: np.ndarray):
    hi = m.astype(np.float16)
    lo = (m - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def tables():
    """Ordered list of (name, [N,16] fp16 logical matrix).  up_lo multiplies only the hi part of x (odd K rows zero):
    x*U ~= (xh + xl)*Uh + xh*Ul."""
    uh, ul = split(up_matrix())
    ul[:, 1::2] = 0
    return [("up_hi", uh), ("up_lo", ul), ("dn_odd", down_matrix(False).astype(np.float16)),
            ("dn_even", down_matrix(True).astype(np.float16))]


def render() -> str:
    words = []
    offs = []
    off = 0
    for name, m in tables():
        phys = swizzle32(m)
        offs.append((name, off, phys.size * 2))
        off += phys.size * 2
        words.append(phys)
    allw = np.concatenate(words).view(np.uint32)
    lines = ["// GENERATED by tools/gen_act_tables.py -- do not edit.  Constant B operands (banded Toeplitz, fp16,",
             "// K-major SWIZZLE_32B) of the tensor-core FIRs in act1d_mma.cu.", "#pragma once", "#include <stdint.h>", ""]
    for name, o, sz in offs:
        lines.append(f"#define HSV_TOEP_{name.upper()} {o}u   // {sz} bytes")
    lines.append(f"#define HSV_TOEP_BYTES {off}u")
    lines.append("")
    lines.append(f"__device__ __align__(16) const uint32_t g_act_toeplitz[{allw.size}] = {{")
    for i in range(0, allw.size, 8):
        lines.append("  " + ", ".join(f"0x{w:08x}u" for w in allw[i:i + 8]) + ",")
    lines.append("};")
    return "\n".join(lines) + "\n"


def main():
    txt = render()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.isfile(OUT) else ""
        if cur != txt:
            print("act_toeplitz_tables.h is stale: run python tools/gen_act_tables.py")
            sys.exit(1)
        print("act_toeplitz_tables.h is up to date")
        return
    with open(OUT, "w") as f:
        f.write(txt)
    print(f"wrote {OUT}")


if __name__ == "__main__":
    main()
