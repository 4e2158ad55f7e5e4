"""Op-for-op torch restatement of the tail of the text-to-vec model (partial): ``W2VDecoder``
(ttv_v1/t2w2v_transformer.py:377-405) and ``PitchPredictor`` (:408-463, ``ResBlock1`` of ttv_v1/modules.py:187-223) --
the last two modules of both ``SynthesizerTrn.infer`` (:1109-1110) and ``inf_plm_gen`` (:991-992); their outputs
``(w2v, pitch)`` are exactly the inputs of ``voice_conversion_noise_control`` (inference.py:158-167).  ORACLE, test
infrastructure (see ``oracle/__init__.py``): state_dict in, tensors out, the same ATen ops in the same order as the
reference modules, pinned bit-exact against them in ``tests/test_oracle_vs_reference.py``."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import functional as OF
from . import functional_front as FF

SD = OF.SD
LRELU_SLOPE = 0.1   # ttv_v1/modules.py:19


def w2v_decoder(sd: SD, p: str, x, x_mask, g, hidden: int = 512, k: int = 5, n_layers: int = 8):
    """W2VDecoder.forward (:399-404)."""
    x = FF._conv(sd, p + "pre.", x * x_mask) * x_mask
    x = FF.wn(sd, p + "enc.", x, x_mask, g, hidden=hidden, k=k, dilation_rate=1, n_layers=n_layers)
    return FF._conv(sd, p + "proj.", x) * x_mask


def resblock1(sd: SD, p: str, x, k: int, dilation=(1, 3, 5)):
    """ResBlock1.forward without a mask (ttv_v1/modules.py:210-223)."""
    for i, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = FF._conv(sd, f"{p}convs1.{i}.", xt, dilation=d, padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = FF._conv(sd, f"{p}convs2.{i}.", xt, padding=(k - 1) // 2)
        x = xt + x
    return x


def pitch_predictor(sd: SD, p: str, x, g, kernels=(3, 5, 7), n_up: int = 2):
    """PitchPredictor.forward (:440-461): [B,1024,T] -> [B,1,4T] (log-f0 at the 4x frame rate)."""
    x = FF._conv(sd, p + "conv_pre.", x, padding=3) + FF._conv(sd, p + "cond.", g)
    nk = len(kernels)
    for i in range(n_up):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, OF.wn_weight(sd, f"{p}ups.{i}."), sd.get(f"{p}ups.{i}.bias"), stride=2, padding=1)
        xs = None
        for j, k in enumerate(kernels):
            r = resblock1(sd, f"{p}resblocks.{i * nk + j}.", x, k)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)                                       # default slope 0.01 (:456)
    return F.conv1d(x, sd[p + "conv_post.weight"], None, padding=3)


def ttv_tail(sd: SD, z, y_mask, g):
    """(:1109-1110) w2v = w2v_decoder(z, y_mask, g); pitch = pp(w2v, g)."""
    w2v = w2v_decoder(sd, "w2v_decoder.", z, y_mask, g)
    return w2v, pitch_predictor(sd, "pp.", w2v, g)
