"""Op-for-op torch-CPU fp32 restatement of the reference hot path (ORACLE).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Every function takes the
reference's ``state_dict`` (same keys) plus a key prefix
and evaluates the same ATen ops in the same order as the reference module it
cites, so on the same machine it agrees with the reference to the last bit.
No nn.Module, no autograd.  All file:line citations are relative to the
reference repository root.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

LRELU_SLOPE = 0.1  # modules.py:17


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """commons.py:14-15."""
    return int((kernel_size * dilation - dilation) / 2)


def wn_weight(sd: SD, p: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = v * (g / ||v||), norm over all
    dims but 0, no epsilon (hierspeechpp_speechsynthesizer.py:401,406)."""
    if p + "weight" in sd:
        return sd[p + "weight"]
    v, g = sd[p + "weight_v"], sd[p + "weight_g"]
    return torch._weight_norm(v, g, 0)


def _bias(sd: SD, p: str):
    return sd.get(p + "bias")


def conv1d(sd: SD, p: str, x, dilation: int = 1, padding: int = 0):
    return F.conv1d(x, wn_weight(sd, p), _bias(sd, p), stride=1, padding=padding, dilation=dilation)


def conv_transpose1d(sd: SD, p: str, x, stride: int, padding: int):
    """ups[i]: hierspeechpp_speechsynthesizer.py:404-408 (weight [Cin,Cout,k], norm per Cin)."""
    return F.conv_transpose1d(x, wn_weight(sd, p), _bias(sd, p), stride=stride, padding=padding)


# --------------------------------------------------------------------------
# alias_free_torch
# --------------------------------------------------------------------------
def upsample1d(x, filt, ratio: int = 2):
    """alias_free_torch/resample.py:11-32."""
    k = filt.shape[-1]
    pad = k // ratio - 1
    pad_left = pad * ratio + (k - ratio) // 2
    pad_right = pad * ratio + (k - ratio + 1) // 2
    C = x.shape[1]
    x = F.pad(x, (pad, pad), mode="replicate")
    x = ratio * F.conv_transpose1d(x, filt.expand(C, -1, -1), stride=ratio, groups=C)
    return x[..., pad_left:-pad_right]


def downsample1d(x, filt, ratio: int = 2):
    """alias_free_torch/resample.py:36-48 -> filter.py:60-94 (LowPassFilter1d)."""
    k = filt.shape[-1]
    even = k % 2 == 0
    pad_left = k // 2 - int(even)
    pad_right = k // 2
    C = x.shape[1]
    x = F.pad(x, (pad_left, pad_right), mode="replicate")
    return F.conv1d(x, filt.expand(C, -1, -1), stride=ratio, groups=C)


def snakebeta(x, alpha, beta):
    """activations.py:107-119 with alpha_logscale=True (every use on the path)."""
    a = torch.exp(alpha.unsqueeze(0).unsqueeze(-1))
    b = torch.exp(beta.unsqueeze(0).unsqueeze(-1))
    return x + (1.0 / (b + 0.000000001)) * torch.pow(torch.sin(x * a), 2)


def activation1d(sd: SD, p: str, x):
    """alias_free_torch/act.py:23-27."""
    x = upsample1d(x, sd[p + "upsample.filter"])
    x = snakebeta(x, sd[p + "act.alpha"], sd[p + "act.beta"])
    return downsample1d(x, sd[p + "downsample.lowpass.filter"])


# --------------------------------------------------------------------------
# AMP blocks, DBlock
# --------------------------------------------------------------------------
def amp_block(sd: SD, p: str, x, kernel_size: int, dilation: Sequence[int] = (1, 3, 5)):
    """AMPBlock1 hierspeechpp_speechsynthesizer.py:377-386 == AMPBlock0 speechsr24k/speechsr.py:49-58."""
    for i, d in enumerate(dilation):
        xt = activation1d(sd, f"{p}activations.{2 * i}.", x)
        xt = conv1d(sd, f"{p}convs1.{i}.", xt, dilation=d, padding=get_padding(kernel_size, d))
        xt = activation1d(sd, f"{p}activations.{2 * i + 1}.", xt)
        xt = conv1d(sd, f"{p}convs2.{i}.", xt, dilation=1, padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def dblock(sd: SD, p: str, x, factor: int = 4):
    """DBlock.forward hierspeechpp_speechsynthesizer.py:328-339."""
    size = x.shape[-1] // factor
    residual = conv1d(sd, p + "residual_dense.", x)
    residual = F.interpolate(residual, size=size)
    x = F.interpolate(x, size=size)
    for i, d in enumerate((1, 2, 4)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = conv1d(sd, f"{p}conv.{i}.", x, dilation=d, padding=d)
    return x + residual


def _mean_of_blocks(sd, p, x, stage, kernel_sizes, dilations):
    nk = len(kernel_sizes)
    xs = None
    for j, (k, d) in enumerate(zip(kernel_sizes, dilations)):
        y = amp_block(sd, f"{p}resblocks.{stage * nk + j}.", x, k, d)
        xs = y if xs is None else xs + y
    return xs / nk


# --------------------------------------------------------------------------
# HierSpeech++ vocoder
# --------------------------------------------------------------------------
def source_network(sd: SD, p: str, x, g):
    """SourceNetwork.forward hierspeechpp_speechsynthesizer.py:290-308 -> (e, e_)."""
    ks, rates, upk = (3, 5, 7), (2, 2), (4, 4)
    dil = ((1, 3, 5),) * 3
    x = conv1d(sd, p + "conv_pre.", x, padding=3) + conv1d(sd, p + "cond.", g)
    for i, (u, k) in enumerate(zip(rates, upk)):
        x = conv_transpose1d(sd, f"{p}ups.{i}.", x, stride=u, padding=(k - u) // 2)
        x = _mean_of_blocks(sd, p, x, i, ks, dil)
    x = activation1d(sd, p + "activation_post.", x)
    x_ = F.conv1d(x, sd[p + "conv_post.weight"], None, padding=3)
    return x, x_


def hier_generator(sd: SD, p: str, x, pitch, g,
                   resblock_kernel_sizes=(3, 7, 11),
                   resblock_dilation_sizes=((1, 3, 5),) * 3,
                   upsample_rates=(4, 5, 4, 2, 2),
                   upsample_kernel_sizes=(8, 11, 8, 4, 4)):
    """Generator.forward hierspeechpp_speechsynthesizer.py:428-451."""
    x = conv1d(sd, p + "conv_pre.", x, padding=3) + dblock(sd, p + "downs.", pitch) + conv1d(sd, p + "cond.", g)
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        x = conv_transpose1d(sd, f"{p}ups.{i}.", x, stride=u, padding=(k - u) // 2)
        if i == 0:
            x = x + conv1d(sd, p + "proj.", pitch, padding=3)
        x = _mean_of_blocks(sd, p, x, i, resblock_kernel_sizes, resblock_dilation_sizes)
    x = activation1d(sd, p + "activation_post.", x)
    x = F.conv1d(x, sd[p + "conv_post.weight"], None, padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------
# SpeechSR
# --------------------------------------------------------------------------
def speechsr_out_len(L: int, which: int) -> int:
    """speechsr24k/speechsr.py:96 ``int(L*1.5)`` / speechsr48k/speechsr.py:96 ``int(L*3)``."""
    return int(L * 1.5) if which == 24 else int(L * 3)


def speechsr_generator(sd: SD, p: str, x, which: int,
                       resblock_kernel_sizes=(3, 7, 11),
                       resblock_dilation_sizes=((1, 3, 5),) * 3):
    """SpeechSR Generator.forward speechsr24k/speechsr.py:89-109 (48k twin :89-109), g=None."""
    x = conv1d(sd, p + "conv_pre.", x, padding=3)
    x = F.interpolate(x, speechsr_out_len(x.shape[-1], which), mode="linear")
    x = _mean_of_blocks(sd, p, x, 0, resblock_kernel_sizes, resblock_dilation_sizes)
    x = activation1d(sd, p + "activation_post.", x)
    x = F.conv1d(x, sd[p + "conv_post.weight"], None, padding=3)
    return torch.tanh(x)


def speechsr(sd: SD, x, which: int):
    """SynthesizerTrn.forward speechsr24k/speechsr.py:244-247."""
    return speechsr_generator(sd, "dec.", x, which)


def vocoder(sd: SD, z, g, sn_prefix="sn.", dec_prefix="dec."):
    """The sn -> dec pair of SynthesizerTrn.infer hierspeechpp_speechsynthesizer.py:648-649."""
    e, _ = source_network(sd, sn_prefix, z, g)
    return hier_generator(sd, dec_prefix, z, e, g)


def peak_norm_pcm16(audio: torch.Tensor, s1: float = 32767.0, s2: float = 0.999) -> "np.ndarray":
    """inference_plm.py:183-188 (s1=32767.0, s2=0.999 or the prompt peak) / inference_speechsr.py:39-41
    (s1=0.999, s2=32767.0): ``audio / abs(audio).max() * s1 * s2`` in fp32, then numpy ``astype('int16')``."""
    a = audio.squeeze()
    a = a / (torch.abs(a).max()) * s1 * s2
    return a.cpu().numpy().astype("int16")
