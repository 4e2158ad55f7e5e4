"""Import shim for the REAL reference — oracle pinning and the reference-vs-B200 tests only.

The reference is a flat script collection whose hot-path modules import a few
third-party packages that are absent here (``timm``, ``monotonic_align``).
``monotonic_align`` is never reached at inference; ``timm``'s ``Attention`` is
reached by the flows of the full ``SynthesizerTrn`` (modules.py:397), so it is
restated functionally below (timm 0.6.13, the version requirements.txt:10 pins).

Roots probed, in order: ``$HSV_REFERENCE_ROOT``, ``/root/reference`` (authoring
container), ``<repo>/baseline/_ref`` (the copy ``baseline/install_ref.py``
makes; git-ignored, it travels to the GPU box).  Callers must check
:func:`available` first.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    cands = [os.environ.get("HSV_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "hierspeechpp_speechsynthesizer.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hierspeechpp_speechsynthesizer.py"))


def _timm_attention():
    """timm 0.6.13 ``timm.models.vision_transformer.Attention`` (requirements.txt:10), restated: qkv Linear ->
    softmax(q k^T / sqrt(head_dim)) v -> proj Linear.  Same parameter names (qkv, proj), so state_dicts match."""
    import torch
    from torch import nn

    class Attention(nn.Module):
        def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
            super().__init__()
            assert dim % num_heads == 0
            self.num_heads = num_heads
            self.scale = (dim // num_heads) ** -0.5
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.attn_drop = nn.Dropout(attn_drop)
            self.proj = nn.Linear(dim, dim)
            self.proj_drop = nn.Dropout(proj_drop)

        def forward(self, x):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            q, k, v = qkv.unbind(0)
            attn = (q @ k.transpose(-2, -1)) * self.scale
            attn = self.attn_drop(attn.softmax(dim=-1))
            x = (attn @ v).transpose(1, 2).reshape(B, N, C)
            return self.proj_drop(self.proj(x))

    return Attention


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Return a namespace with the reference's hot-path modules imported."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    import transformers  # noqa: F401  (must precede the timm stub)

    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.vision_transformer", Attention=_timm_attention())
    _stub("monotonic_align", mask_from_lens=None)
    _stub("monotonic_align.core", maximum_path_c=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging

    lvl = logging.getLogger().level
    utils = importlib.import_module("utils")
    logging.getLogger().setLevel(max(lvl, logging.WARNING))
    H = importlib.import_module("hierspeechpp_speechsynthesizer")
    sr24 = importlib.import_module("speechsr24k.speechsr")
    sr48 = importlib.import_module("speechsr48k.speechsr")
    aft = importlib.import_module("alias_free_torch")
    act = importlib.import_module("activations")
    _loaded.update(utils=utils, H=H, sr24=sr24, sr48=sr48, alias_free_torch=aft, activations=act)
    return types.SimpleNamespace(**_loaded)


def load_ttv():
    """The reference's text-to-vec module (``ttv_v1/t2w2v_transformer.py``) -- used for ``W2VDecoder`` (:377-405) and
    ``PitchPredictor`` (:408-463).  Its import chain pulls training-only packages that this image lacks
    (``torchmetrics``, ``matplotlib``); they are stubbed, none of them is on the inference path of the two classes."""
    load()
    if "ttv" in _loaded:
        return _loaded["ttv"]

    class _Unused:
        def __init__(self, *a, **k):
            pass

    _stub("torchmetrics")
    _stub("torchmetrics.classification", MulticlassAccuracy=_Unused)
    _stub("matplotlib")
    _stub("matplotlib.pyplot", Figure=_Unused)      # only a type annotation in ttv_v1/utils_mega.py:41
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    ma = sys.modules["monotonic_align"]
    if not hasattr(ma, "mask_from_lens") or ma.mask_from_lens is None:
        ma.mask_from_lens = _Unused
    _loaded["ttv"] = importlib.import_module("ttv_v1.t2w2v_transformer")
    return _loaded["ttv"]


W2V_DECODER_CFG = dict(in_channels=256, hidden_channels=512, kernel_size=5, dilation_rate=1, n_layers=8,
                       output_size=1024, p_dropout=0.1, gin_channels=256)   # ttv_v1/t2w2v_transformer.py (TTV decoder)


HIER_SYNTH_CFG = dict(spec_channels=513, segment_size=30, inter_channels=192, hidden_channels=192,
                      filter_channels=768, n_heads=2, n_layers=6, kernel_size=3, p_dropout=0.1, resblock="1",
                      resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
                      upsample_rates=[4, 5, 4, 2, 2], upsample_initial_channel=512,
                      upsample_kernel_sizes=[8, 11, 8, 4, 4], gin_channels=256)


def build_synthesizer(seed: int = 1234):
    """The reference's own ``SynthesizerTrn`` (hierspeechpp_speechsynthesizer.py:562-633) with the libritts960
    architecture of SURVEY.md §B.1, seeded random init (no checkpoint is shipped), SnakeBeta
    parameters drawn as in ``synth.vocoder_sd`` so that the activation is exercised.  eval mode, CPU."""
    import torch

    ref = load()
    torch.manual_seed(seed)
    m = ref.H.SynthesizerTrn(**HIER_SYNTH_CFG)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p_ in m.named_parameters():
            if n.endswith("act.alpha"):
                p_.copy_(torch.rand(p_.shape, generator=gen) * 1.5 - 0.5)
            elif n.endswith("act.beta"):
                p_.copy_(torch.rand(p_.shape, generator=gen) * 1.3 - 0.5)
            elif n.endswith("post.weight") or n.endswith("adaLN_modulation.1.weight"):
                # the reference zero-initialises these (flows = identity at init): give the flows something to do
                p_.copy_(torch.randn(p_.shape, generator=gen) * 0.02)
    m.eval()
    return m


def load_speechsr(which: int):
    """Reference SpeechSR (24 or 48) with its bundled checkpoint, eval mode, CPU."""
    ref = load()
    import torch

    d = os.path.join(REFERENCE_ROOT, f"speechsr{which}k")
    ckpt = os.path.join(d, "G_340000.pth" if which == 24 else "G_100000.pth")
    h = ref.utils.get_hparams_from_file(os.path.join(d, "config.json"))
    cls = ref.sr24.SynthesizerTrn if which == 24 else ref.sr48.SynthesizerTrn
    m = cls(h.data.n_mel_channels, h.train.segment_size // h.data.hop_length, **h.model)
    sd = torch.load(ckpt, map_location="cpu")["model"]
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def example_wav():
    """example/reference_1.wav as float32 [1,1,48000] in [-1,1)."""
    import numpy as np
    import torch
    from scipy.io import wavfile

    sr, w = wavfile.read(os.path.join(REFERENCE_ROOT, "example", "reference_1.wav"))
    assert sr == 16000 and w.dtype == np.int16
    return torch.from_numpy(w.astype(np.float32) / 32768.0).view(1, 1, -1)
