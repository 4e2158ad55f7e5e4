"""Harmonic sine source (north_star item 3): the 64-bit fixed-point phase accumulator does not drift over 30 s."""
import numpy as np
import pytest
import torch

from oracle import closed_form as CF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _f0(B, T, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(T)[None, :]
    f0 = 180.0 + 120.0 * np.sin(2 * np.pi * t / 173.0 + rng.uniform(0, 6.28, (B, 1))) + rng.uniform(-20, 20, (B, T))
    f0[rng.uniform(size=(B, T)) < 0.25] = 0.0            # unvoiced frames
    return f0.astype(np.float32)


@pytest.mark.parametrize("B,T,hop,H", [(2, 1500, 320, 8), (1, 7, 80, 3), (3, 500, 320, 1)])
def test_sinegen_vs_fp64(hsv, B, T, hop, H):
    sr = 16000.0
    f0 = _f0(B, T, B * 100 + T)
    ref, uv_ref = CF.sinegen(f0, hop, sr, H, 0.1)
    out, uv = hsv.ops.sinegen(torch.from_numpy(f0).to(DEV), hop, sr, H, 0.1)
    out, uv = out.cpu().numpy(), uv.cpu().numpy()
    assert out.shape == ref.shape and np.array_equal(uv, uv_ref)
    # phase error budget: 2^-24 turns of the final fp32 angle x harmonic number (+ sinpif's ~1e-7); no growth with n
    err = np.abs(out - ref) / 0.1
    tol = 2 * np.pi * 2.0 ** -24 * np.arange(1, H + 1)[None, :, None] + 3e-7
    assert np.all(err <= tol), float((err / tol).max())
    tail = err[:, :, -hop * 5:]                           # the last 100 ms of the 30 s utterance: no drift
    assert tail.max() <= (2 * np.pi * 2.0 ** -24 * H + 3e-7)


def test_sinegen_fp32_cumsum_would_drift(hsv):
    """What the accumulator avoids: a float32 running sum of the same increments is off by > 1e-3 rad after 30 s."""
    sr, hop, T = 16000.0, 320, 1500
    f0 = _f0(1, T, 5)
    up = np.repeat(f0, hop, axis=1)
    naive = np.cumsum(np.where(up > 0, up / np.float32(sr), 0).astype(np.float32), axis=1, dtype=np.float32)
    exact = np.cumsum(np.where(up > 0, up.astype(np.float64) / sr, 0.0), axis=1)
    drift_naive = np.abs(naive.astype(np.float64) - exact)[0, -1] * 2 * np.pi
    ref, _ = CF.sinegen(f0, hop, sr, 1, 1.0)
    out, _ = hsv.ops.sinegen(torch.from_numpy(f0).to(DEV), hop, sr, 1, 1.0)
    err_end = np.abs(out.cpu().numpy() - ref)[0, 0, -2000:].max()
    assert drift_naive > 1e-3 and err_end < 1e-4, (drift_naive, err_end)


def test_sinegen_is_deterministic_and_batch_independent(hsv):
    f0 = torch.from_numpy(_f0(3, 400, 9)).to(DEV)
    a, _ = hsv.ops.sinegen(f0, 320, 16000.0, 4)
    b, _ = hsv.ops.sinegen(f0, 320, 16000.0, 4)
    c, _ = hsv.ops.sinegen(f0[1:2].contiguous(), 320, 16000.0, 4)
    assert torch.equal(a, b) and torch.equal(a[1:2], c)
