"""The reference's OWN modules, patched, against themselves unpatched -- on the GPU.

``hsv.patch_reference()`` swaps the hot-path classes inside the imported reference modules; the reference's
``SynthesizerTrn`` (hierspeechpp_speechsynthesizer.py:562-633) / SpeechSR ``SynthesizerTrn`` (speechsr.py:215-252) is
then built by the reference's own constructor, loads the unpatched model's ``state_dict`` strictly, and its own
``voice_conversion_noise_control`` / ``infer`` / ``forward`` (:675-699, :635-651, speechsr.py:244-252) is compared with
the unpatched reference running eagerly in strict fp32 on the same device.  Needs the reference sources: ``/root/reference``
here, ``baseline/_ref`` (baseline/install_ref.py) on the GPU box; skipped when neither exists."""
import contextlib
import importlib
import sys

import numpy as np
import pytest
import torch

from conftest import MAX_ABS_TOL, SNR_DB_MIN, golden
from oracle import closed_form as CF
from oracle import refload

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refload.available(), reason="reference sources not present")]
DEV = "cuda:0"


@contextlib.contextmanager
def patched(hsv, ref):
    names = {"H": ("Generator", "SourceNetwork", "AMPBlock1", "DBlock", "Activation1d", "PosteriorSFEncoder",
                   "ResidualCouplingBlock_Transformer", "StyleEncoder"),
             "sr24": ("Generator", "AMPBlock0", "Activation1d"), "sr48": ("Generator", "AMPBlock0", "Activation1d")}
    saved = {(mod, n): getattr(getattr(ref, mod), n) for mod, ns in names.items() for n in ns}
    try:
        yield hsv.patch_reference()
    finally:
        for (mod, n), v in saved.items():
            setattr(getattr(ref, mod), n, v)
        for modname in ("alias_free_torch", "activations"):
            importlib.reload(sys.modules[modname])


@contextlib.contextmanager
def strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _check(name, got, ref, max_abs=MAX_ABS_TOL):
    got, ref = got.detach().float().cpu().numpy(), ref.detach().float().cpu().numpy()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    ma, snr = CF.max_abs(ref, got), CF.snr_db(ref, got)
    print(f"[parity] {name}: max_abs={ma:.3e} snr={snr:.1f} dB")
    assert ma <= max_abs and snr >= SNR_DB_MIN, (name, ma, snr)


def test_reference_synthesizer_with_patched_vocoder(hsv):
    ref = refload.load()
    A = refload.build_synthesizer(1234)                   # the reference, untouched
    state = A.state_dict()
    with patched(hsv, ref) as names:
        assert "hierspeechpp_speechsynthesizer.Generator" in names
        torch.manual_seed(1234)
        B = ref.H.SynthesizerTrn(**refload.HIER_SYNTH_CFG)   # the reference's constructor builds the B200 classes
        assert isinstance(B.dec, hsv.Generator) and isinstance(B.sn, hsv.SourceNetwork)
        assert isinstance(B.enc_p_l, hsv.front.PosteriorSFEncoder) and isinstance(B.emb_g, hsv.front.StyleEncoder)
        assert isinstance(B.flow, hsv.front.ResidualCouplingBlock_Transformer)
        B.load_state_dict(state, strict=True)
        B.eval()
    A.to(DEV); B.to(DEV)
    T = 150                                                # 3 s
    gen = torch.Generator().manual_seed(1111)
    w2v = torch.randn(1, 1024, T, generator=gen).to(DEV)
    hz = torch.rand(1, 1, 4 * T, generator=gen) * 320 + 80
    hz[torch.rand(1, 1, 4 * T, generator=gen) < 0.3] = 0.0
    f0 = torch.log(hz + 1).to(DEV)
    mel = (torch.randn(2, 80, 150, generator=gen) * 2 - 4).to(DEV)
    ln, ln2 = torch.LongTensor([T]).to(DEV), torch.LongTensor([150, 150]).to(DEV)
    with strict_fp32(), torch.no_grad():
        torch.manual_seed(7)                               # the randn_like of :687
        oa = A.voice_conversion_noise_control(w2v, ln, mel, ln2, f0, noise_scale=0.333, denoise_ratio=0.3)
        torch.manual_seed(7)
        ob = B.voice_conversion_noise_control(w2v, ln, mel, ln2, f0, noise_scale=0.333, denoise_ratio=0.3)
        assert ob.shape == (1, 1, 320 * T)
        _check("reference SynthesizerTrn.voice_conversion_noise_control, patched vs unpatched", ob, oa)
        torch.manual_seed(8)                               # enc_p_l samples z = m + randn * exp(logs) (:198-200)
        oa2, ea = A.infer(mel[:1], w2v, torch.LongTensor([150]).to(DEV), f0)
        torch.manual_seed(8)
        ob2, eb = B.infer(mel[:1], w2v, torch.LongTensor([150]).to(DEV), f0)
        _check("reference SynthesizerTrn.infer wav", ob2, oa2)
        _check("reference SynthesizerTrn.infer e_ (predicted f0)", eb, ea,
               max_abs=MAX_ABS_TOL * max(1.0, float(ea.abs().max())))


@pytest.mark.parametrize("which", [24, 48])
def test_reference_speechsr_patched(hsv, which):
    ref = refload.load()
    A = refload.load_speechsr(which)                       # bundled checkpoint, unpatched
    mod = ref.sr24 if which == 24 else ref.sr48
    h = ref.utils.get_hparams_from_file(f"{refload.REFERENCE_ROOT}/speechsr{which}k/config.json")
    with patched(hsv, ref):
        B = mod.SynthesizerTrn(h.data.n_mel_channels, h.train.segment_size // h.data.hop_length, **h.model)
        assert isinstance(B.dec, hsv.SpeechSRGenerator)
        B.load_state_dict(A.state_dict(), strict=True)
        B.eval()
    A.to(DEV); B.to(DEV)
    x = refload.example_wav().to(DEV)                       # the full 3 s example
    with strict_fp32(), torch.no_grad():
        ya, yb = A(x), B(x)
        _check(f"reference speechsr{which}k SynthesizerTrn.forward, patched vs unpatched (3 s example)", yb, ya)
        _check(f"reference speechsr{which}k infer(max_len)", B.infer(x, max_len=16000), A.infer(x, max_len=16000))
    if which == 24:                                         # config #1 known answer (CPU reference output)
        g = golden("speechsr24_example.npz")
        _check("speechsr24k patched vs CPU golden", yb, torch.from_numpy(g["y"]))
