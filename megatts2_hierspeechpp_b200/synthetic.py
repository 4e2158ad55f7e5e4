"""Seeded synthetic checkpoints and inputs (random-init weights of the reference's architectures).

The HierSpeech++ vocoder checkpoint is not shipped with the reference
(SURVEY.md §0.4), so config #2/#5 run on random-init weights.  This module
builds ``state_dict``s with exactly the reference's keys and shapes
 from a seed, without importing the reference, so the
same weights exist in the authoring container (where they are loaded
``strict=True`` into the real reference modules to pin the oracle) and on the
GPU box (where only the oracle and the CUDA path exist).

Init follows PyTorch's Conv1d default (U(+-1/sqrt(fan_in)) for weight_v and
bias, weight_g = ||v||), SnakeBeta alpha ~ U(-0.5, 1.0), beta ~ U(-0.5, 0.8)
(the 'tame' ranges of SURVEY.md §7.3 / §8d).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

import numpy as np

from .modules import FILTER_TAPS

FILTER_TAPS_F32 = np.array(FILTER_TAPS, dtype=np.float32)  # the fp32 taps stored in the reference checkpoints

SD = Dict[str, torch.Tensor]


def _conv(sd: SD, p: str, gen, cout: int, cin: int, k: int, wn: bool = True, bias: bool = True,
          transposed: bool = False):
    # Conv1d weight [Cout,Cin,k]; ConvTranspose1d weight [Cin,Cout,k] (fan_in = dim1*k in torch's rule)
    shape = (cin, cout, k) if transposed else (cout, cin, k)
    fan_in = shape[1] * k
    bound = 1.0 / math.sqrt(fan_in)
    v = (torch.rand(shape, generator=gen) * 2 - 1) * bound
    if wn:
        sd[p + "weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1).clone()
        sd[p + "weight_v"] = v
    else:
        sd[p + "weight"] = v
    if bias:
        sd[p + "bias"] = (torch.rand(cout, generator=gen) * 2 - 1) * bound


def _act(sd: SD, p: str, gen, c: int, tame: bool = True):
    lo_a, hi_a, lo_b, hi_b = (-0.5, 1.0, -0.5, 0.8) if tame else (-0.95, 2.41, -2.93, 0.80)
    sd[p + "act.alpha"] = torch.rand(c, generator=gen) * (hi_a - lo_a) + lo_a
    sd[p + "act.beta"] = torch.rand(c, generator=gen) * (hi_b - lo_b) + lo_b
    f = torch.from_numpy(FILTER_TAPS_F32.copy()).view(1, 1, 12)
    sd[p + "upsample.filter"] = f.clone()
    sd[p + "downsample.lowpass.filter"] = f.clone()


def _amp_block(sd: SD, p: str, gen, c: int, k: int, tame: bool = True):
    for i in range(3):
        _conv(sd, f"{p}convs1.{i}.", gen, c, c, k)
    for i in range(3):
        _conv(sd, f"{p}convs2.{i}.", gen, c, c, k)
    for i in range(6):
        _act(sd, f"{p}activations.{i}.", gen, c, tame)


HIER_CFG = dict(initial_channel=192, resblock_kernel_sizes=[3, 7, 11],
                resblock_dilation_sizes=[[1, 3, 5]] * 3, upsample_rates=[4, 5, 4, 2, 2],
                upsample_initial_channel=512, upsample_kernel_sizes=[8, 11, 8, 4, 4], gin_channels=256)

SR_CFG = dict(resblock="0", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
              upsample_rates=[3], upsample_initial_channel=32, upsample_kernel_sizes=[3])


def hier_generator_sd(seed: int = 1234, prefix: str = "dec.", cfg=None, tame: bool = True) -> SD:
    """Keys of hierspeechpp_speechsynthesizer.Generator (:395-426)."""
    cfg = cfg or HIER_CFG
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    c0 = cfg["upsample_initial_channel"]
    _conv(sd, prefix + "conv_pre.", gen, c0, cfg["initial_channel"], 7)
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        _conv(sd, f"{prefix}ups.{i}.", gen, c0 // 2 ** (i + 1), c0 // 2 ** i, k, transposed=True)
    nk = len(cfg["resblock_kernel_sizes"])
    for i in range(len(cfg["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, k in enumerate(cfg["resblock_kernel_sizes"]):
            _amp_block(sd, f"{prefix}resblocks.{i * nk + j}.", gen, ch, k, tame)
    _act(sd, prefix + "activation_post.", gen, ch, tame)
    _conv(sd, prefix + "conv_post.", gen, 1, ch, 7, wn=False, bias=False)
    _conv(sd, prefix + "cond.", gen, c0, cfg["gin_channels"], 1, wn=False)
    _conv(sd, prefix + "downs.residual_dense.", gen, c0, c0 // 8, 1)
    _conv(sd, prefix + "downs.conv.0.", gen, c0, c0 // 8, 3)
    _conv(sd, prefix + "downs.conv.1.", gen, c0, c0, 3)
    _conv(sd, prefix + "downs.conv.2.", gen, c0, c0, 3)
    _conv(sd, prefix + "proj.", gen, c0 // 2, c0 // 8, 7, wn=False)
    return sd


def source_network_sd(seed: int = 1235, prefix: str = "sn.", c0: int = 256, tame: bool = True) -> SD:
    """Keys of hierspeechpp_speechsynthesizer.SourceNetwork (:252-287)."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv(sd, prefix + "conv_pre.", gen, c0, 192, 7)
    for i in range(2):
        _conv(sd, f"{prefix}ups.{i}.", gen, c0 // 2 ** (i + 1), c0 // 2 ** i, 4, transposed=True)
    ch = c0
    for i in range(2):
        ch = c0 // 2 ** (i + 1)
        for j, k in enumerate((3, 5, 7)):
            _amp_block(sd, f"{prefix}resblocks.{i * 3 + j}.", gen, ch, k, tame)
    _act(sd, prefix + "activation_post.", gen, ch, tame)
    _conv(sd, prefix + "conv_post.", gen, 1, ch, 7, wn=False, bias=False)
    _conv(sd, prefix + "cond.", gen, c0, 256, 1, wn=False)
    return sd


def vocoder_sd(seed: int = 1234) -> SD:
    sd = hier_generator_sd(seed, "dec.")
    sd.update(source_network_sd(seed + 1, "sn."))
    return sd


def speechsr_sd(seed: int = 4321, prefix: str = "dec.", tame: bool = True) -> SD:
    """Keys of speechsr24k/speechsr.py Generator (:67-87): 134 tensors."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv(sd, prefix + "conv_pre.", gen, 32, 1, 7)
    for j, k in enumerate((3, 7, 11)):
        _amp_block(sd, f"{prefix}resblocks.{j}.", gen, 32, k, tame)
    _act(sd, prefix + "activation_post.", gen, 32, tame)
    _conv(sd, prefix + "conv_post.", gen, 1, 32, 7, wn=False, bias=False)
    return sd


def vocoder_inputs(B: int, T: int, seed: int = 1111):
    """z ~ N(0,1) [B,192,T], g ~ N(0,1) [B,256,1]."""
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(B, 192, T, generator=gen)
    g = torch.randn(B, 256, 1, generator=gen)
    return z, g


def speechsr_input(B: int, L: int, seed: int = 1111):
    """x = 0.1*N(0,1) [B,1,L]."""
    gen = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(B, 1, L, generator=gen)


# ----------------------------------------------------------------------------------------------
# the step BEFORE the vocoder: enc_p_l, flow_l / flow, emb_g of SynthesizerTrn
# ----------------------------------------------------------------------------------------------
def _linear(sd: SD, p: str, gen, cout: int, cin: int, scale: float = None):
    bound = 1.0 / math.sqrt(cin) if scale is None else scale
    sd[p + "weight"] = (torch.rand(cout, cin, generator=gen) * 2 - 1) * bound
    sd[p + "bias"] = (torch.rand(cout, generator=gen) * 2 - 1) * bound


def _wn(sd: SD, p: str, gen, hidden: int, k: int, n_layers: int, gin: int):
    """modules.WN (:111-182): weight-normed cond_layer, in_layers, res_skip_layers."""
    for i in range(n_layers):
        _conv(sd, f"{p}in_layers.{i}.", gen, 2 * hidden, hidden, k)
    for i in range(n_layers):
        _conv(sd, f"{p}res_skip_layers.{i}.", gen, 2 * hidden if i < n_layers - 1 else hidden, hidden, 1)
    _conv(sd, p + "cond_layer.", gen, 2 * hidden * n_layers, gin, 1)


def posterior_sf_encoder_sd(seed: int, prefix: str = "enc_p_l.", hidden: int = 192, out: int = 192, k: int = 5,
                            n_layers: int = 16, gin: int = 256, src: int = 1024) -> SD:
    """Keys of hierspeechpp_speechsynthesizer.PosteriorSFEncoder (:168-203)."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv(sd, prefix + "pre_source.", gen, hidden, src, 1, wn=False)
    _conv(sd, prefix + "pre_filter.", gen, hidden, 1, 9, wn=False)
    for name in ("source_enc.", "filter_enc.", "enc."):
        _wn(sd, prefix + name, gen, hidden, k, n_layers // 2, gin)
    _conv(sd, prefix + "proj.", gen, 2 * out, hidden, 1, wn=False)
    return _ordered_like_reference(sd, prefix)


def _ordered_like_reference(sd: SD, prefix: str) -> SD:
    """weight-normed convs register bias, weight_g, weight_v in that order (torch.nn.utils.weight_norm)."""
    out: SD = {}
    done = set()
    for key in sd:
        base = key.rsplit(".", 1)[0] + "."
        if base in done:
            continue
        if base + "weight_g" in sd:
            for leaf in ("bias", "weight_g", "weight_v"):
                if base + leaf in sd:
                    out[base + leaf] = sd[base + leaf]
        else:
            for leaf in ("weight", "bias"):
                if base + leaf in sd:
                    out[base + leaf] = sd[base + leaf]
        done.add(base)
    return out


def coupling_block_sd(seed: int, prefix: str = "flow_l.", channels: int = 192, hidden: int = 192, n_layers: int = 3,
                      n_flows: int = 4, gin: int = 256) -> SD:
    """Keys of ResidualCouplingBlock_Transformer (:53-88) with ResidualCouplingLayer_Transformer_simple flows
    (modules.py:412-488; DiTConVBlock :390-411).  The reference zero-initialises ``post`` and the adaLN output layer
    (identity flows); here they get small random values so that the flows do something."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _linear(sd, prefix + "cond_block.0.", gen, 4 * hidden, gin)
    _linear(sd, prefix + "cond_block.2.", gen, hidden, 4 * hidden)
    half = channels // 2
    for f in range(n_flows):
        p = f"{prefix}flows.{2 * f}."
        _conv(sd, p + "pre.", gen, hidden, half, 1, wn=False)
        for b in range(n_layers):
            q = f"{p}enc_block.{b}."
            _linear(sd, q + "attn.qkv.", gen, 3 * hidden, hidden)
            _linear(sd, q + "attn.proj.", gen, hidden, hidden)
            _conv(sd, q + "mlp.fc1.", gen, 4 * hidden, hidden, 5, wn=False)
            _conv(sd, q + "mlp.fc2.", gen, hidden, 4 * hidden, 1, wn=False)
            _linear(sd, q + "adaLN_modulation.1.", gen, 6 * hidden, hidden, scale=0.02)
        _conv(sd, p + "post.", gen, half, hidden, 1, wn=False)
        sd[p + "post.weight"] *= 0.3
    return sd


def style_encoder_sd(seed: int, prefix: str = "emb_g.", in_dim: int = 80, hidden: int = 256, out: int = 256) -> SD:
    """Keys of styleencoder.StyleEncoder (:33-66)."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv(sd, prefix + "spectral.0.", gen, hidden, in_dim, 1, wn=False)
    _conv(sd, prefix + "spectral.3.", gen, hidden, hidden, 1, wn=False)
    for i in range(2):
        _conv(sd, f"{prefix}temporal.{i}.conv1.", gen, 2 * hidden, hidden, 5, wn=False)
    for name in ("conv_q.", "conv_k.", "conv_v.", "conv_o."):
        _conv(sd, prefix + "slf_attn." + name, gen, hidden, hidden, 1, wn=False)
    _conv(sd, prefix + "fc.", gen, out, hidden, 1, wn=False)
    return sd


def front_sd(seed: int = 2345) -> SD:
    """enc_p_l + flow_l + flow + emb_g of the HierSpeech++ SynthesizerTrn (libritts960 architecture)."""
    sd = posterior_sf_encoder_sd(seed, "enc_p_l.")
    sd.update(coupling_block_sd(seed + 1, "flow_l."))
    sd.update(coupling_block_sd(seed + 2, "flow."))
    sd.update(style_encoder_sd(seed + 3, "emb_g."))
    return sd


def synthesizer_sd(seed: int = 1234) -> SD:
    """Everything SynthesizerTrn.infer / voice_conversion_noise_control touches: front + sn + dec."""
    sd = front_sd(seed + 1111)
    sd.update(vocoder_sd(seed))
    return sd


def synthesizer_inputs(T: int, T_mel: int = 150, seed: int = 1111):
    """SURVEY.md §8d #2, SynthesizerTrn level: w2v ~ N(0,1) [1,1024,T], f0 = log(hz+1) with hz ~ U(80,400) and 30 %
    unvoiced [1,1,4T], trg_mel ~ N(-4,2) [2,80,T_mel]."""
    gen = torch.Generator().manual_seed(seed)
    w2v = torch.randn(1, 1024, T, generator=gen)
    hz = torch.rand(1, 1, 4 * T, generator=gen) * 320 + 80
    hz[torch.rand(1, 1, 4 * T, generator=gen) < 0.3] = 0.0
    f0 = torch.log(hz + 1)
    mel = torch.randn(2, 80, T_mel, generator=gen) * 2 - 4
    return w2v, f0, mel


# ----------------------------------------------------------------------------------------------
# the tail of the text-to-vec model (partial): W2VDecoder + PitchPredictor
# ----------------------------------------------------------------------------------------------
def w2v_decoder_sd(seed: int = 3456, prefix: str = "w2v_decoder.", cin: int = 256, hidden: int = 512, k: int = 5,
                   n_layers: int = 8, out: int = 1024, gin: int = 256) -> SD:
    """Keys of ttv_v1/t2w2v_transformer.W2VDecoder (:377-405) as instantiated at :779."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv(sd, prefix + "pre.", gen, hidden, cin, 1, wn=False)
    _wn(sd, prefix + "enc.", gen, hidden, k, n_layers, gin)
    _conv(sd, prefix + "proj.", gen, out, hidden, 1, wn=False)
    return _ordered_like_reference(sd, prefix)


def pitch_predictor_sd(seed: int = 3457, prefix: str = "pp.", cin: int = 1024, c0: int = 256, gin: int = 256,
                       kernels=(3, 5, 7)) -> SD:
    """Keys of ttv_v1/t2w2v_transformer.PitchPredictor (:408-438): conv_pre, 2 weight-normed ConvTranspose1d(k4, u2),
    2 x 3 ResBlock1, conv_post (no bias), cond."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv(sd, prefix + "conv_pre.", gen, c0, cin, 7, wn=False)
    for i in range(2):
        _conv(sd, f"{prefix}ups.{i}.", gen, c0 >> (i + 1), c0 >> i, 4, transposed=True)
    for i in range(2):
        ch = c0 >> (i + 1)
        for j, k in enumerate(kernels):
            for grp in ("convs1", "convs2"):
                for l in range(3):
                    _conv(sd, f"{prefix}resblocks.{i * len(kernels) + j}.{grp}.{l}.", gen, ch, ch, k)
    _conv(sd, prefix + "conv_post.", gen, 1, c0 >> 2, 7, wn=False, bias=False)
    _conv(sd, prefix + "cond.", gen, c0, gin, 1, wn=False)
    return _ordered_like_reference(sd, prefix)


def ttv_tail_sd(seed: int = 3456) -> SD:
    sd = w2v_decoder_sd(seed)
    sd.update(pitch_predictor_sd(seed + 1))
    return sd


def ttv_tail_inputs(B: int, T: int, seed: int = 1111, lengths=None):
    """(z [B,256,T], y_mask [B,1,T], g [B,256,1]): the flow output, its frame mask and the style vector."""
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(B, 256, T, generator=gen)
    g = torch.randn(B, 256, 1, generator=gen) * 0.5
    ln = torch.full((B,), T, dtype=torch.long) if lengths is None else torch.as_tensor(lengths, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < ln[:, None]).unsqueeze(1).to(torch.float32)
    return z, mask, g
