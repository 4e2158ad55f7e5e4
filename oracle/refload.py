"""Import shim for the REAL reference (``/root/reference``) — oracle pinning only.

The reference is a flat script collection whose hot-path modules import a few
third-party packages that are absent here (``timm``, ``monotonic_align``).
Neither is reached by the waveform-generation path, so placeholders are
enough.  ``/root/reference`` does not exist on the GPU
box; callers must check :func:`available` first.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HSV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hierspeechpp_speechsynthesizer.py"))


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

    class _Attention:  # never called on the hot path
        def __init__(self, *a, **k):
            raise RuntimeError("timm Attention placeholder")

    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.vision_transformer", Attention=_Attention)
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
