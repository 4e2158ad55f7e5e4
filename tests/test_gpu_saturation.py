"""fp16 operand health counter (hsv_blk16_stats): the reference is fp32 end to end, the tensor-core operands here are
fp16, so a checkpoint whose activations exceed +-65504 must be detectable."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_blk16_stats_counts_saturation(hsv):
    ops = hsv.ops
    B, C, L = 2, 32, 700
    x = torch.randn(B, C, L, device=DEV)
    x[1, 5, 100] = 7.0e4       # beyond fp16
    x[0, 31, 699] = -1.0e5
    x[0, 0, 0] = 1234.5
    buf = ops.blk16_buffer(B, C, L, DEV, slot=9)
    ops.pack_blk16(x, buf)
    st = ops.blk16_stats(buf, C, L).cpu()
    assert int(st[0]) == 2
    assert abs(float(st[1:2].view(torch.float32)) - 1234.5) <= 1.0
    x2 = torch.randn(B, C, L, device=DEV)
    ops.pack_blk16(x2, buf)
    st = ops.blk16_stats(buf, C, L).cpu()
    assert int(st[0]) == 0 and float(st[1:2].view(torch.float32)) == float(x2.half().abs().max())


def test_module_level_saturation_report(hsv):
    from oracle import synth
    ops = hsv.ops
    m = hsv.Vocoder()
    m.load_state_dict(synth.vocoder_sd(1234), strict=True)
    m.to(DEV).eval()
    z, g = synth.vocoder_inputs(1, 20, seed=1111)
    ops.SATURATION["enabled"] = True
    try:
        with torch.no_grad():
            m(z.to(DEV), g.to(DEV))
        n, mx = ops.saturation_report()
        assert n == 0 and 0.0 < mx < 6.0e4          # seeded net: operands far inside the fp16 range
        print(f"[parity] fp16 operand peak over one vocoder forward: {mx:.3f} (0 non-finite)")
    finally:
        ops.SATURATION["enabled"] = False
