"""Op-for-op torch restatement of the step BEFORE the vocoder: ``StyleEncoder``,
``PosteriorSFEncoder`` and the reverse pass of the two ``ResidualCouplingBlock_Transformer`` flows, i.e. everything
``SynthesizerTrn.infer`` / ``voice_conversion_noise_control`` (hierspeechpp_speechsynthesizer.py:635-699) runs before
``sn`` / ``dec``.  ORACLE, test infrastructure (see ``oracle/__init__.py``): state_dict in, tensors out, the same ATen
ops in the same order as the reference modules cited, pinned bit-exact against them in
``tests/test_oracle_vs_reference.py``.  File:line citations are relative to the reference root.  ``timm``'s Attention
(0.6.13, requirements.txt:10) is restated from its published forward."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import functional as OF

SD = OF.SD


def sequence_mask(length, max_length=None):
    """commons.py:128-132."""
    if max_length is None:
        max_length = length.max()
    x = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def _conv(sd: SD, p: str, x, **kw):
    return F.conv1d(x, OF.wn_weight(sd, p), sd.get(p + "bias"), **kw)


def wn(sd: SD, p: str, x, x_mask, g, hidden: int = 192, k: int = 5, dilation_rate: int = 1, n_layers: int = 8):
    """modules.WN.forward (modules.py:147-174)."""
    output = torch.zeros_like(x)
    g = _conv(sd, p + "cond_layer.", g)
    for i in range(n_layers):
        d = dilation_rate ** i
        x_in = _conv(sd, f"{p}in_layers.{i}.", x, dilation=d, padding=int((k * d - d) / 2))
        g_l = g[:, i * 2 * hidden:(i + 1) * 2 * hidden, :]
        in_act = x_in + g_l                                   # commons.fused_add_tanh_sigmoid_multiply (:108-114)
        acts = torch.tanh(in_act[:, :hidden, :]) * torch.sigmoid(in_act[:, hidden:, :])
        rs = _conv(sd, f"{p}res_skip_layers.{i}.", acts)
        if i < n_layers - 1:
            x = (x + rs[:, :hidden, :]) * x_mask
            output = output + rs[:, hidden:, :]
        else:
            output = output + rs
    return output * x_mask


def posterior_sf_encoder(sd: SD, p: str, x_src, x_ftr, x_mask, g, eps=None, out: int = 192):
    """PosteriorSFEncoder.forward (hierspeechpp_speechsynthesizer.py:192-203) -> (z, m, logs)."""
    xs = _conv(sd, p + "pre_source.", x_src) * x_mask
    xf = _conv(sd, p + "pre_filter.", x_ftr, stride=4, padding=4) * x_mask
    xs = wn(sd, p + "source_enc.", xs, x_mask, g)
    xf = wn(sd, p + "filter_enc.", xf, x_mask, g)
    x = wn(sd, p + "enc.", xs + xf, x_mask, g)
    stats = _conv(sd, p + "proj.", x) * x_mask
    m, logs = torch.split(stats, out, dim=1)
    if eps is None:
        eps = torch.randn_like(m)
    z = (m + eps * torch.exp(logs)) * x_mask
    return z, m, logs


def timm_attention(sd: SD, p: str, x, num_heads: int = 2):
    """timm 0.6.13 vision_transformer.Attention.forward; x [B,N,C]."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = (q @ k.transpose(-2, -1)) * ((C // num_heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def _modulate(x, shift, scale):
    """modules.py:346-347."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def dit_conv_block(sd: SD, p: str, x, c, x_mask, hidden: int = 192):
    """modules.DiTConVBlock.forward (:405-410); x [B,T,C], c [B,C], x_mask [B,T,1]."""
    x = x * x_mask
    mod = F.linear(F.silu(c), sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    h = _modulate(F.layer_norm(x, (hidden,), None, None, 1e-6) * x_mask, shift_msa, scale_msa)
    x = x + gate_msa.unsqueeze(1) * timm_attention(sd, p + "attn.", h) * x_mask
    h = _modulate(F.layer_norm(x, (hidden,), None, None, 1e-6), shift_mlp, scale_mlp)
    # FFN_Conv.forward (:382-388), mask given as [B,1,T]
    m2 = x_mask.transpose(1, 2)
    y = F.conv1d(h.transpose(1, 2), sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"], padding=2)
    y = F.gelu(y, approximate="tanh")
    y = F.conv1d(y * m2, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]) * m2
    return x + gate_mlp.unsqueeze(1) * y.transpose(1, 2)


def coupling_layer_reverse(sd: SD, p: str, x, x_mask, g, half: int = 96, n_layers: int = 3):
    """ResidualCouplingLayer_Transformer_simple.forward(reverse=True), mean_only (modules.py:455-488)."""
    x0, x1 = torch.split(x, [half] * 2, 1)
    h = _conv(sd, p + "pre.", x0) * x_mask
    h = h.transpose(1, 2)
    mt = x_mask.transpose(1, 2)
    for b in range(n_layers):
        h = dit_conv_block(sd, f"{p}enc_block.{b}.", h, g, mt)
    h = h.transpose(1, 2)
    m = _conv(sd, p + "post.", h) * x_mask
    logs = torch.zeros_like(m)
    x1 = (x1 - m) * torch.exp(-logs) * x_mask
    return torch.cat([x0, x1], 1)


def coupling_block_reverse(sd: SD, p: str, x, x_mask, g, n_flows: int = 4):
    """ResidualCouplingBlock_Transformer.forward(reverse=True) (:78-88): for flow in reversed([C0,F,C1,F,..])."""
    c = F.linear(g.squeeze(2), sd[p + "cond_block.0.weight"], sd[p + "cond_block.0.bias"])
    c = F.linear(F.silu(c), sd[p + "cond_block.2.weight"], sd[p + "cond_block.2.bias"])
    for f in reversed(range(n_flows)):
        x = torch.flip(x, [1])                                # modules.Flip (:270-277)
        x = coupling_layer_reverse(sd, f"{p}flows.{2 * f}.", x, x_mask, c)
    return x


def multi_head_attention(sd: SD, p: str, x, c, attn_mask, n_heads: int = 2):
    """attentions.MultiHeadAttention.forward / attention (:147-188) without relative / proximal terms."""
    q, k, v = _conv(sd, p + "conv_q.", x), _conv(sd, p + "conv_k.", c), _conv(sd, p + "conv_v.", c)
    b, d, t_s, t_t = (*k.size(), q.size(2))
    kc = d // n_heads
    q = q.view(b, n_heads, kc, t_t).transpose(2, 3)
    k = k.view(b, n_heads, kc, t_s).transpose(2, 3)
    v = v.view(b, n_heads, kc, t_s).transpose(2, 3)
    scores = torch.matmul(q / math.sqrt(kc), k.transpose(-2, -1))
    if attn_mask is not None:
        scores = scores.masked_fill(attn_mask == 0, -1e4)
    p_attn = F.softmax(scores, dim=-1)
    out = torch.matmul(p_attn, v).transpose(2, 3).contiguous().view(b, d, t_t)
    return _conv(sd, p + "conv_o.", out)


def _mish(x):
    return x * torch.tanh(F.softplus(x))


def style_encoder(sd: SD, p: str, x, mask):
    """StyleEncoder.forward (styleencoder.py:68-89) -> [B, out_dim]."""
    x = _mish(_conv(sd, p + "spectral.0.", x))
    x = _mish(_conv(sd, p + "spectral.3.", x)) * mask
    for i in range(2):                                        # Conv1dGLU (:20-31)
        y = _conv(sd, f"{p}temporal.{i}.conv1.", x, padding=2)
        y1, y2 = torch.split(y, y.shape[1] // 2, dim=1)
        x = x + y1 * torch.sigmoid(y2)
    x = x * mask
    attn_mask = mask.unsqueeze(2) * mask.unsqueeze(-1)
    x = x + multi_head_attention(sd, p + "slf_attn.", x, x, attn_mask)
    x = _conv(sd, p + "fc.", x)
    return torch.div(x.sum(dim=2), mask.sum(dim=2))           # temporal_avg_pool (:91-99)


def front(sd: SD, w2v, src_length, trg_mel, trg_length, f0, noise_scale: float = 0.333, denoise_ratio: float = 0.0,
          eps=None):
    """voice_conversion_noise_control up to z (hierspeechpp_speechsynthesizer.py:677-690) -> (z, g)."""
    trg_mask = torch.unsqueeze(sequence_mask(trg_length, trg_mel.size(2)), 1).to(trg_mel.dtype)
    g = style_encoder(sd, "emb_g.", trg_mel, trg_mask).unsqueeze(-1)
    g = (1 - denoise_ratio) * g[:1] + denoise_ratio * g[1:]
    y_mask = torch.unsqueeze(sequence_mask(src_length, w2v.size(2)), 1).to(trg_mel.dtype)
    # the reference draws TWO noise tensors: inside enc_p_l (:201, discarded) and at :687
    _, m_p, logs_p = posterior_sf_encoder(sd, "enc_p_l.", w2v, f0, y_mask, g)
    if eps is None:
        eps = torch.randn_like(m_p)
    z = (m_p + eps * torch.exp(logs_p) * noise_scale) * y_mask
    z = coupling_block_reverse(sd, "flow_l.", z, y_mask, g)
    z = coupling_block_reverse(sd, "flow.", z, y_mask, g)
    return z, g


def voice_conversion_noise_control(sd: SD, w2v, src_length, trg_mel, trg_length, f0, noise_scale: float = 0.333,
                                   denoise_ratio: float = 0.0, eps=None):
    """SynthesizerTrn.voice_conversion_noise_control (:675-699)."""
    z, g = front(sd, w2v, src_length, trg_mel, trg_length, f0, noise_scale, denoise_ratio, eps)
    e, _ = OF.source_network(sd, "sn.", z, g)
    return OF.hier_generator(sd, "dec.", z, e, g)
