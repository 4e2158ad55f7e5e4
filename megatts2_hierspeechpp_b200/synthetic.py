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
