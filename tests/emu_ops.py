"""CPU emulation of the C-ABI ops for HOST-LOGIC tests only (tests/test_host_dataflow.py).

It lets the drop-in modules' dataflow (buffer reuse, epilogue modes, residual order, weight caches)
run on CPU against the oracle without a GPU.  It is test infrastructure: never imported by the
package, never measured."""
import torch
import torch.nn.functional as F

from megatts2_hierspeechpp_b200 import ops as real
from oracle import functional as OF
from oracle.closed_form import FILTER_TAPS_F32

_T = torch.from_numpy(FILTER_TAPS_F32.copy()).view(1, 1, 12)
_pool = {}


def _act(x, a, b):
    sd = {"p.act.alpha": a, "p.act.beta": b, "p.upsample.filter": _T, "p.downsample.lowpass.filter": _T}
    return OF.activation1d(sd, "p.", x)


def blk16_buffer(B, C, L, device, slot=0):
    key = (B, C, L, slot)
    if key not in _pool:
        _pool[key] = torch.zeros(B, C // 8, real.blk16_rows(L), 8, dtype=torch.float16)
    return _pool[key]


def _unpack(buf, L):
    B, nch, Lp, _ = buf.shape
    return buf[:, :, real.BLK_PAD:real.BLK_PAD + L, :].float().permute(0, 1, 3, 2).reshape(B, nch * 8, L)


def _pack_into(buf, y):
    B, C, L = y.shape
    buf[:, :, real.BLK_PAD:real.BLK_PAD + L, :] = y.half().view(B, C // 8, 8, L).permute(0, 1, 3, 2)


def act1d(x, a, b, out=None, scale=1.0):
    y = _act(x * scale, a, b)
    if out is not None:
        out.copy_(y)
        return out
    return y


def act1d_blk16(x, a, b, buf, scale=1.0):
    _pack_into(buf, _act(x * scale, a, b))
    return buf


def pack_blk16(x, buf, lrelu=False, scale=1.0):
    if isinstance(x, (list, tuple)):
        xs = list(x)
        x = xs[0]
        for t in xs[1:]:
            x = x + t
    x = x * scale
    _pack_into(buf, F.leaky_relu(x, 0.1) if lrelu else x)
    return buf


def weight_norm_fold(v, g):
    return torch._weight_norm(v, g, 0)


def pack_conv_weight(w, n_tile):
    p = w.half().flatten().clone()
    p._emu_shape = tuple(w.shape)
    return p


def conv1d_umma(a_blk, wp, bias, L, cin, cout, k, d, n_tile, residual=None, out=None, acc=None,
                acc_mode=0, acc_div=1.0, want_out=True):
    x = _unpack(a_blk, L)
    w = wp.float().view(cout, cin, k)
    v = F.conv1d(x, w, bias, padding=(k - 1) // 2 * d, dilation=d)
    if residual is not None:
        v = v + residual
    if acc_mode == 1:
        acc.copy_(v)
    elif acc_mode == 2:
        acc.add_(v)
    if out is not None:
        out.copy_(v)
        return out
    return v if want_out else None


def conv1d_umma_blk(a_blk, wp, bias, L, cin, cout, k, d, n_tile, out_blk, mode=0, bc=None, mask=None):
    x = _unpack(a_blk, L)
    w = wp.float().view(cout, cin, k)
    v = F.conv1d(x, w, bias, padding=(k - 1) // 2 * d, dilation=d)
    if bc is not None:
        v = v + bc.unsqueeze(-1)
    if mode == real.BLK_GATE:
        B = v.shape[0]
        g = v.view(B, cout // 16, 2, 8, L)
        v = (torch.tanh(g[:, :, 0]) * torch.sigmoid(g[:, :, 1])).reshape(B, cout // 2, L)
    elif mode == real.BLK_GELU:
        v = F.gelu(v, approximate="tanh")
    elif mode == real.BLK_LRELU:
        v = F.leaky_relu(v, 0.1)
    _pack_into(out_blk, v * _m(mask, v))
    return out_blk


def conv1d_umma_wn_tail(a_blk, wp, bias, x, output, mask, next_blk, n_tile):
    B, H, L = x.shape
    cin = a_blk.shape[1] * a_blk.shape[3]
    rs = F.conv1d(_unpack(a_blk, L), wp.float().view(2 * H, cin, 1), bias)
    return wn_res_pack(x, rs, mask, output, next_blk)


def act_conv1d_umma(x, alpha, beta, wp, bias, cout, k, d, residual=None, out=None, acc=None, acc_mode=0, scale=1.0,
                    want_out=True):
    B, cin, L = x.shape
    assert out is None or out.data_ptr() != x.data_ptr()
    assert acc is None or acc.data_ptr() != x.data_ptr()
    a = _act(x * scale, alpha, beta).half().float()          # the kernel's fp16 operand rounding
    v = F.conv1d(a, wp.float().view(cout, cin, k), bias, padding=(k - 1) // 2 * d, dilation=d)
    if residual is not None:
        v = v + residual
    if acc_mode == 1:
        acc.copy_(v)
    elif acc_mode == 2:
        acc.add_(v)
    if out is not None:
        out.copy_(v)
        return out
    return v if want_out else None


def pack_convT_weight(w, u, n_tile):
    return w.half().flatten().clone()


def conv_transpose1d_umma(a_blk, wp, bias, Lin, cin, cout, k, u, n_tile, add=None):
    x = _unpack(a_blk, Lin)
    v = F.conv_transpose1d(x, wp.float().view(cin, cout, k), bias, stride=u, padding=(k - u) // 2)
    return v if add is None else v + add


def conv1d_direct(x, w, bias, d=1, pad=0, flags=0, out=None):
    xin = F.leaky_relu(x, 0.1) if flags & real.CONV_LRELU_IN else x
    if flags & real.CONV_LRELU001_IN:
        xin = F.leaky_relu(xin, 0.01)
    if flags & real.CONV_SILU_IN:
        xin = F.silu(xin)
    v = F.conv1d(xin, w, bias, padding=pad, dilation=d)
    if flags & real.CONV_TANH:
        v = torch.tanh(v)
    if flags & real.CONV_ADD_OUT:
        out.add_(v)
        return out
    if out is not None:
        out.copy_(v)
        return out
    return v


def conv_transpose1d(x, w, bias, u, add=None):
    k = w.shape[-1]
    v = F.conv_transpose1d(x, w, bias, stride=u, padding=(k - u) // 2)
    return v if add is None else v + add


def sr_pre_interp(x, w, bias, Lout):
    return F.interpolate(F.conv1d(x, w, bias, padding=3), Lout, mode="linear")


def nearest_gather(x, Lout):
    return F.interpolate(x, size=Lout)


def add3_bcast(a, b, bc, out=None):
    v = a if b is None else a + b
    if bc is not None:
        v = v + bc.view(a.shape[0], a.shape[1], 1)
    if out is not None:
        out.copy_(v)
        return out
    return v


# ---- frame-rate operators (csrc/frame_ops.cu) ----
def _m(mask, like):
    return 1.0 if mask is None else mask.view(mask.shape[0], 1, -1)


def pack_blk16_act(x, buf, C, mode=0, bcast=None, mask=None, c_off=0):
    mk = _m(mask, x)
    x = x[:, c_off:]
    if mode == real.PACK_GATE:
        b = 0.0 if bcast is None else bcast.view(bcast.shape[0], -1, 1)
        a = x[:, :2 * C] + b
        y = torch.tanh(a[:, :C]) * torch.sigmoid(a[:, C:])
    elif mode == real.PACK_GELU:
        y = F.gelu(x[:, :C], approximate="tanh") * mk
    elif mode == real.PACK_MISH:
        y = x[:, :C] * torch.tanh(F.softplus(x[:, :C])) * mk
    else:
        y = x[:, :C] * mk
    _pack_into(buf, y.contiguous())
    return buf


def wn_res_pack(x, rs, mask, output, buf):
    C = x.shape[1]
    x.copy_((x + rs[:, :C]) * _m(mask, x))
    output.add_(rs[:, C:])
    _pack_into(buf, x)
    return buf


def ln_mod_blk16(x, shift, scale, buf, mod_stride, mask=None, eps=1e-6, inmask=False, premask=False):
    mk = _m(mask, x)
    if inmask:
        x = x * mk
    n = F.layer_norm(x.transpose(1, 2), (x.shape[1],), None, None, eps).transpose(1, 2)
    if premask:
        n = n * mk
    _pack_into(buf, (n * (1 + scale.unsqueeze(-1)) + shift.unsqueeze(-1)).contiguous())
    return buf


def gate_ln_mod_blk16(x, y, gate, shift, scale, buf, mod_stride, mask=None, eps=1e-6, premask=False):
    x.add_(gate.unsqueeze(-1) * y * _m(mask, x))
    return ln_mod_blk16(x, shift, scale, buf, mod_stride, mask, eps, premask=premask)


def frame_op(op, a, b=None, c=None, mask=None, out=None, out2=None, B=0, C=0, L=0, s=1.0, cstride=0):
    mk = _m(mask, a if a is not None else b)
    if op == real.OP_WN_RES:
        res, skip = b[:, :C], b[:, C:]
        out.copy_((a + res) * mk); out2.add_(skip)
    elif op == real.OP_WN_LAST:
        out2.copy_((out2 + b) * mk)
    elif op == real.OP_GATE_ADD:
        out.copy_(a + (c.unsqueeze(-1) * b) * mk)
    elif op == real.OP_COUPLE:
        out[:, C:2 * C] = (a[:, C:2 * C] - b) * mk
    elif op == real.OP_SAMPLE:
        out.copy_((a[:, :C] + (b * torch.exp(a[:, C:2 * C])) * s) * mk)
    elif op == real.OP_MASK:
        out.copy_(a * mk)
    elif op == real.OP_ADD:
        out.copy_((a + b) * mk)
    elif op == real.OP_GLU_RES:
        out.copy_((a + b[:, :C] * torch.sigmoid(b[:, C:])) * mk)
    elif op == real.OP_MISH:
        out.copy_(a * torch.tanh(F.softplus(a)) * mk)
    elif op == real.OP_FLIP:
        out.copy_(torch.flip(a, [1]))
    elif op == real.OP_ADD_BCAST:
        out.copy_((a + c.unsqueeze(-1)) * mk)
    else:
        raise ValueError(op)


def mha(q, k, v, B, heads, D, Tq, Tk, q_bs, k_bs, v_bs, scale, prescale_q, lens=None, out_blk=None):
    def grab(t, bs, T):
        return torch.stack([t.reshape(-1)[b * bs:b * bs + heads * D * T].view(heads, D, T) for b in range(B)])
    qq, kk, vv = grab(q, q_bs, Tq).transpose(2, 3), grab(k, k_bs, Tk).transpose(2, 3), grab(v, v_bs, Tk).transpose(2, 3)
    sc = torch.matmul(qq * scale, kk.transpose(-2, -1)) if prescale_q else torch.matmul(qq, kk.transpose(-2, -1)) * scale
    if lens is not None:
        for b in range(B):
            n = int(lens[b])
            sc[b, :, n:, :] = -1e4
            sc[b, :, :, n:] = -1e4
    o = torch.matmul(sc.softmax(-1), vv)
    o = o.transpose(2, 3).contiguous().view(B, heads * D, Tq)
    if out_blk is not None:
        _pack_into(out_blk, o)
        return out_blk
    return o


def conv1d_c1_strided(x, w, bias, stride, pad, mask=None):
    return F.conv1d(x, w, bias, stride=stride, padding=pad) * _m(mask, x)


def masked_mean(x, mask):
    den = float(x.shape[-1]) if mask is None else mask.sum(dim=1, keepdim=True)
    return x.sum(dim=2) / den


def check_saturation(buf, C, L):
    return None


NAMES = ["pack_blk16_act", "wn_res_pack", "ln_mod_blk16", "gate_ln_mod_blk16", "frame_op", "mha", "conv1d_c1_strided", "masked_mean", "check_saturation",
         "blk16_buffer", "act1d", "act1d_blk16", "pack_blk16", "weight_norm_fold", "pack_conv_weight", "conv1d_umma", "conv1d_umma_blk", "conv1d_umma_wn_tail", "act_conv1d_umma", "pack_convT_weight", "conv_transpose1d_umma",
         "conv1d_direct", "conv_transpose1d", "sr_pre_interp", "nearest_gather", "add3_bcast"]


def install(monkeypatch):
    import megatts2_hierspeechpp_b200.modules as M
    for n in NAMES:
        monkeypatch.setattr(real, n, globals()[n])
    monkeypatch.setattr(M, "_as_input", lambda x: x.detach().contiguous())
    _pool.clear()
