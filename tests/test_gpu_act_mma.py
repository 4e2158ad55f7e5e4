"""GPU parity of the tensor-core activation kernel (hsv_act1d_snakebeta, out_mode 1, variant 2 = act1d_mma.cu)
against the oracle's Activation1d (alias_free_torch/act.py:23-27 restated in oracle/functional.py)."""
import numpy as np
import pytest
import torch

from oracle import closed_form as CF
from oracle import functional as OF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle(x, alpha, beta):
    t = torch.from_numpy(CF.FILTER_TAPS_F32.copy()).view(1, 1, 12)
    sd = {"a.act.alpha": alpha, "a.act.beta": beta, "a.upsample.filter": t, "a.downsample.lowpass.filter": t}
    return OF.activation1d(sd, "a.", x)


@pytest.fixture
def mma(hsv):
    from megatts2_hierspeechpp_b200 import _lib
    lib = _lib.load()
    lib.hsv_set_act_variant(2)     # force the tensor-core variant for every shape
    yield hsv
    lib.hsv_set_act_variant(0)


def _run(hsv, x, al, be, scale=1.0, slot=6):
    B, C, L = x.shape
    buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=slot)
    hsv.ops.act1d_blk16(x.to(DEV), al.to(DEV), be.to(DEV), buf, scale=scale)
    y = hsv.ops.unpack_blk16(buf, C, L).cpu()
    from megatts2_hierspeechpp_b200 import ops
    assert buf[:, :, :ops.BLK_PAD].abs().max().item() == 0 and buf[:, :, ops.BLK_PAD + L:].abs().max().item() == 0
    return y


# tolerance: fp16 rounding of the 2x-rate signal (2^-11, 12 taps) + fp16 rounding of the result (2^-11)
def _check(y, ref, what=""):
    tol = 1e-3 * max(1.0, ref.abs().max().item())
    err = (y - ref).abs().max().item()
    assert err <= tol, (what, err, tol)


@pytest.mark.parametrize("B,C,L", [(1, 16, 448), (1, 16, 449), (1, 16, 512), (2, 32, 1000), (1, 64, 2000), (1, 16, 1),
                                   (1, 16, 2), (1, 16, 5), (1, 16, 31), (1, 32, 33), (1, 16, 447), (1, 16, 895),
                                   (1, 16, 897), (2, 128, 700), (1, 96, 300), (1, 256, 2000), (1, 16, 1001),
                                   (3, 48, 1343), (1, 16, 4099)])
def test_act_mma_vs_oracle(mma, B, C, L):
    gen = torch.Generator().manual_seed(B * 7919 + C * 31 + L)
    x = torch.randn(B, C, L, generator=gen) * 1.5
    al = torch.rand(C, generator=gen) * 1.5 - 0.5
    be = torch.rand(C, generator=gen) * 1.3 - 0.5
    _check(_run(mma, x, al, be), _oracle(x, al, be), (B, C, L))


def test_act_mma_checkpoint_range_params(mma):
    """alpha in [-1, 2.4], beta in [-2.9, 0.8]: the range of the bundled SpeechSR checkpoints (sine amplified ~19x)."""
    gen = torch.Generator().manual_seed(5)
    C, L = 32, 3000
    x = torch.randn(2, C, L, generator=gen) * 2
    al = torch.rand(C, generator=gen) * 3.4 - 1.0
    be = torch.rand(C, generator=gen) * 3.7 - 2.9
    _check(_run(mma, x, al, be), _oracle(x, al, be))


def test_act_mma_in_scale(mma):
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(2, 16, 700, generator=gen) * 3
    al = torch.rand(16, generator=gen) - 0.5
    be = torch.rand(16, generator=gen) - 0.5
    _check(_run(mma, x, al, be, scale=1.0 / 3), _oracle(x / 3, al, be))


def test_act_mma_matches_cuda_core_kernel(hsv):
    """Both variants of the same entry point agree to the fp16 rounding of z."""
    from megatts2_hierspeechpp_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(3)
    B, C, L = 2, 64, 5000
    x = torch.randn(B, C, L, generator=gen).to(DEV)
    al = (torch.rand(C, generator=gen) - 0.5).to(DEV)
    be = (torch.rand(C, generator=gen) - 0.5).to(DEV)
    outs = []
    try:
        for v in (1, 2):
            lib.hsv_set_act_variant(v)
            buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=6 + v)
            hsv.ops.act1d_blk16(x, al, be, buf)
            outs.append(hsv.ops.unpack_blk16(buf, C, L))
    finally:
        lib.hsv_set_act_variant(0)
    assert (outs[0] - outs[1]).abs().max().item() <= 1e-3 * max(1.0, outs[0].abs().max().item())


def test_act_mma_full_size_properties(mma):
    """Config-#2 stage-4 size [2,16,160000] and a SpeechSR48 layer slice [1,32,480000]: determinism, utterance
    independence, window spot checks against the oracle (interior, both ends)."""
    hsv = mma
    for (B, C, L) in ((2, 16, 160000), (1, 32, 480000)):
        gen = torch.Generator().manual_seed(C)
        x = torch.randn(B, C, L, generator=gen)
        al = torch.rand(C, generator=gen) - 0.5
        be = torch.rand(C, generator=gen) - 0.5
        y = _run(hsv, x, al, be)
        assert torch.equal(y, _run(hsv, x, al, be, slot=5))                    # deterministic
        if B > 1:
            assert torch.equal(y[1:], _run(hsv, x[1:].contiguous(), al, be, slot=4))   # utterances independent
        for sl in (slice(0, 700), slice(L // 2 - 300, L // 2 + 400), slice(L - 700, L)):
            xw = x[:1, :, max(0, sl.start - 16):min(L, sl.stop + 16)]
            ref = _oracle(xw, al, be)
            o0 = sl.start - max(0, sl.start - 16)
            got = y[:1, :, sl]
            refw = ref[:, :, o0:o0 + got.shape[-1]]
            inner = slice(0 if sl.start == 0 else 8, None if sl.stop == L else -8)
            _check(got[:, :, inner], refw[:, :, inner], (B, C, L, sl))
