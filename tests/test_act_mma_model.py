"""CPU pins for the tensor-core activation kernel (csrc/act1d_mma.cu): the committed Toeplitz-table header is what
tools/gen_act_tables.py generates, and the numpy model of the kernel's tiling / K-blocks / output windows / hi-lo
splits (tests/emu_act_mma.py) agrees with the fp64 closed form of Activation1d (oracle/closed_form.py, SURVEY.md
§A.1) to the fp16 rounding of the 2x-rate signal -- including the replicate-clamp edges and ragged lengths."""
import os
import subprocess
import sys

import numpy as np
import pytest

import emu_act_mma as E
from oracle import closed_form as CF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_toeplitz_header_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_act_tables.py"), "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_toeplitz_tables_reproduce_the_taps():
    import gen_act_tables as T
    f = CF.FILTER_TAPS_F32.astype(np.float64)
    g = T.TAPS16.astype(np.float64)
    # fp16 taps: symmetric, within 1.5 fp16 ulp of the fp32 taps, both polyphase sums exactly 0.5 (unit DC gain)
    assert np.array_equal(g, g[::-1])
    assert np.all(np.abs(g - f) <= 1.5 * np.spacing(np.abs(f).astype(np.float16)).astype(np.float64))
    assert g[0::2].sum() == 0.5 and g[1::2].sum() == 0.5
    tabs = dict(T.tables())
    # up-FIR taps: hi + lo reproduces the fp32 coefficient to 2^-22; the lo part of x meets only the hi taps
    up = T.up_matrix().astype(np.float64)
    got = tabs["up_hi"].astype(np.float64) + tabs["up_lo"].astype(np.float64)
    assert np.abs(got[:, 0::2] - up[:, 0::2]).max() <= 2.0 ** -22
    assert np.abs(tabs["up_lo"][:, 1::2]).max() == 0
    for name in ("dn_odd", "dn_even"):
        assert set(np.unique(tabs[name].astype(np.float64))) <= set(np.concatenate([g, [0.0]]))


@pytest.mark.parametrize("L", [1, 2, 5, 31, 32, 33, 447, 448, 449, 512, 896, 1000])
def test_model_vs_closed_form(L):
    rng = np.random.default_rng(L)
    C = 2
    x = (rng.standard_normal((1, C, L)) * 2).astype(np.float32)
    al = rng.uniform(-1.0, 2.4, C).astype(np.float32)      # the bundled checkpoints' range
    be = rng.uniform(-2.9, 0.8, C).astype(np.float32)
    ref = CF.activation1d(x, al, be)
    got = E.activation1d(x, al, be)
    # error budget: z rounded to fp16 (2^-11 relative per sample, 12 taps with sum f^2 = 0.43)
    # (+ the fp16 taps of the low-pass FIR: <= 1.2 ulp, 2e-4 of the white-noise part of z)
    assert np.abs(got - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max())


def test_model_in_scale():
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((1, 1, 700)) * 3).astype(np.float32)
    al, be = np.array([0.3], np.float32), np.array([-0.2], np.float32)
    ref = CF.activation1d(x / np.float32(3), al, be)
    got = E.activation1d(x, al, be, sc=1.0 / 3)
    assert np.abs(got - ref).max() <= 6e-4 * max(1.0, np.abs(ref).max())
