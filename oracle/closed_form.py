"""Independent numpy closed forms for the hot path (ORACLE, test infrastructure).

These restate the *mathematics* of the reference ops
without calling the ops themselves, so they cross-check ``functional`` (which
replays the ATen ops) and give exact integer index tables for the bit-exact
part of the parity contract.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------
# kaiser-sinc filter  (alias_free_torch/filter.py:28-57)
# --------------------------------------------------------------------------
def kaiser_sinc_filter1d(cutoff: float = 0.25, half_width: float = 0.3, kernel_size: int = 12) -> np.ndarray:
    """fp64 evaluation of the 12-tap filter; cast to fp32 by the caller."""
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    n = np.arange(kernel_size, dtype=np.float64)
    # torch.kaiser_window(periodic=False): I0(beta*sqrt(1-((n-(N-1)/2)/((N-1)/2))^2)) / I0(beta)
    r = (n - (kernel_size - 1) / 2.0) / ((kernel_size - 1) / 2.0)
    window = np.i0(beta * np.sqrt(np.clip(1.0 - r * r, 0.0, None))) / np.i0(beta)
    if kernel_size % 2 == 0:
        time = np.arange(-half_size, half_size, dtype=np.float64) + 0.5
    else:
        time = np.arange(kernel_size, dtype=np.float64) - half_size
    filt = 2 * cutoff * window * np.sinc(2 * cutoff * time)
    return filt / filt.sum()


# the fp32 taps as stored in both bundled checkpoints
FILTER_TAPS_F32 = np.array(
    [0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971,
     0.4432097971, 0.1285725832, -0.0576573834, -0.0255434588, 0.0093894657, 0.0020289647],
    dtype=np.float32)


# --------------------------------------------------------------------------
# Activation1d closed form
# --------------------------------------------------------------------------
def up_indices(L: int) -> np.ndarray:
    """Integer table [2L, 6]: source index of tap i for 2x sample n (replicate clamp).

    n even: x[clamp(n/2 + 2 - i)] with taps f[2i+1]; n odd: x[clamp((n+1)/2 + 2 - i)] with f[2i].
    Derived from resample.py:25-32 (pad 5, conv_transpose stride 2, crop 15)."""
    n = np.arange(2 * L)[:, None]
    i = np.arange(6)[None, :]
    m = np.where(n % 2 == 0, n // 2, (n + 1) // 2)
    return np.clip(m + 2 - i, 0, L - 1)


def down_indices(L: int) -> np.ndarray:
    """Integer table [L, 12]: 2x-rate source index of tap j for output t
    (filter.py:77-93: pad_left 5, pad_right 6, stride 2)."""
    t = np.arange(L)[:, None]
    j = np.arange(12)[None, :]
    return np.clip(2 * t + j - 5, 0, 2 * L - 1)


def activation1d(x: np.ndarray, alpha: np.ndarray, beta: np.ndarray, f: np.ndarray | None = None) -> np.ndarray:
    """fp64 closed form of act.py:23-27 for x [B,C,L]; alpha,beta log-scale [C]."""
    f = FILTER_TAPS_F32.astype(np.float64) if f is None else np.asarray(f, np.float64)
    x = np.asarray(x, np.float64)
    B, C, L = x.shape
    ui = up_indices(L)
    n = np.arange(2 * L)
    taps = np.where((n % 2 == 0)[:, None], f[1::2][None, :], f[0::2][None, :])  # [2L,6]
    y = 2.0 * np.einsum("bcni,ni->bcn", x[:, :, ui], taps)
    a = np.exp(np.asarray(alpha, np.float64))[None, :, None]
    ib = 1.0 / (np.exp(np.asarray(beta, np.float64)) + 1e-9)[None, :, None]
    z = y + ib * np.sin(y * a) ** 2
    di = down_indices(L)
    return np.einsum("bctj,j->bct", z[:, :, di], f)


# --------------------------------------------------------------------------
# Interpolation index tables
# --------------------------------------------------------------------------
def linear_interp_table(L_in: int, L_out: int, fma: bool = False):
    """(i0, i1, lam) of F.interpolate(mode='linear', align_corners=False, size=L_out),
    evaluated in fp32 like ATen (``scale*(dst+0.5)-0.5`` clamped at 0).

    ``fma=False`` is the CPU kernel (separate multiply and subtract rounding);
    ``fma=True`` is a single-rounding fused multiply-add as nvcc emits for the CUDA kernel."""
    scale = np.float32(np.float32(L_in) / np.float32(L_out))
    dst = np.arange(L_out, dtype=np.float32) + np.float32(0.5)
    if fma:
        src = (scale.astype(np.float64) * dst.astype(np.float64) - 0.5).astype(np.float32)
    else:
        src = (scale * dst).astype(np.float32) - np.float32(0.5)
    src = np.maximum(src, np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), L_in - 1)
    i1 = i0 + (i0 < L_in - 1)
    lam = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, lam


def nearest_index(L_in: int, L_out: int) -> np.ndarray:
    """F.interpolate(size=L_out) default 'nearest': idx = min(floor(dst*scale), L_in-1), scale fp32."""
    scale = np.float32(np.float32(L_in) / np.float32(L_out))
    idx = np.floor(np.arange(L_out, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, L_in - 1)


# --------------------------------------------------------------------------
# ConvTranspose1d polyphase table
# --------------------------------------------------------------------------
def conv_transpose_phase_table(k: int, u: int):
    """For output phase rho = o mod u: list of (tap j, input offset c-i) with
    o = u*q + rho reading input row q + (c - i) through tap j = r + i*u,
    r = (rho+p) mod u, c = (rho+p) div u, p = (k-u)//2."""
    p = (k - u) // 2
    table = []
    for rho in range(u):
        r, c = (rho + p) % u, (rho + p) // u
        table.append([(j, c - i) for i, j in enumerate(range(r, k, u))])
    return table


def conv_transpose1d(x: np.ndarray, w: np.ndarray, b, u: int) -> np.ndarray:
    """fp64 scatter form: out[co,o] = b + sum x[ci,t] W[ci,co,j], o = t*u + j - p."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    B, Cin, L = x.shape
    _, Cout, k = w.shape
    p = (k - u) // 2
    full = np.zeros((B, Cout, (L - 1) * u + k))
    for j in range(k):
        full[:, :, j:j + (L - 1) * u + 1:u] += np.einsum("bit,io->bot", x, w[:, :, j])
    out = full[:, :, p:p + u * L]
    if b is not None:
        out = out + np.asarray(b, np.float64)[None, :, None]
    return out


def conv1d(x: np.ndarray, w: np.ndarray, b, d: int) -> np.ndarray:
    """fp64 'same' dilated conv: out[co,t] = b + sum W[co,ci,j] x[ci, t+(j-(k-1)/2)d], zero pad."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    B, Cin, L = x.shape
    Cout, _, k = w.shape
    h = (k - 1) // 2 * d
    xp = np.pad(x, ((0, 0), (0, 0), (h, h)))
    out = np.zeros((B, Cout, L))
    for j in range(k):
        out += np.einsum("bit,oi->bot", xp[:, :, j * d:j * d + L], w[:, :, j])
    if b is not None:
        out = out + np.asarray(b, np.float64)[None, :, None]
    return out


# --------------------------------------------------------------------------
# parity metrics (north_star: max-abs <= 2e-3, SNR >= 40 dB)
# --------------------------------------------------------------------------
def max_abs(a, b) -> float:
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def snr_db(ref, test) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(test, np.float64) - ref
    den = float(np.sum(err * err))
    if den == 0.0:
        return float("inf")
    return 10.0 * math.log10(float(np.sum(ref * ref)) / den)


# --------------------------------------------------------------------------
# Harmonic sine source (BASELINE.json north_star item 3; no counterpart in the reference)
# --------------------------------------------------------------------------
def sinegen(f0: np.ndarray, hop: int, sample_rate: float, harmonics: int = 8, amp: float = 0.1):
    """fp64 closed form: phase[n] = sum_{k<=n} f0_up[k]/sr (turns, unvoiced frames hold the phase),
    out[h] = amp * voiced * sin(2 pi (h+1) phase).  f0 [B,T] Hz -> (out [B,H,T*hop] fp64, uv [B,1,T*hop])."""
    f0 = np.asarray(f0, np.float32)
    up = np.repeat(f0, hop, axis=1).astype(np.float64)
    voiced = up > 0
    inc = np.where(voiced, up / float(sample_rate), 0.0)
    phase = np.cumsum(inc, axis=1)                       # fp64: 53 bits, error ~1e-10 turns over 5e5 samples
    phase -= np.floor(phase)
    h = np.arange(1, harmonics + 1, dtype=np.float64)[None, :, None]
    out = amp * voiced[:, None, :] * np.sin(2.0 * np.pi * h * phase[:, None, :])
    return out, voiced[:, None, :].astype(np.float32)
