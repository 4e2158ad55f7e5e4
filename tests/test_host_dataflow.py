"""Host-side dataflow of the drop-in modules, run on CPU with the kernels EMULATED (tests/emu_ops.py).

Checks what the Python layer is responsible for: op order, epilogue modes (residual / mean over
resblocks), buffer reuse, DBlock's gather-before-1x1 rewrite, weight-cache invalidation.  The real
kernels are checked on the GPU (tests/test_gpu_*.py)."""
import numpy as np
import torch

import emu_ops
from conftest import golden, golden_sd
import megatts2_hierspeechpp_b200 as hsv
from oracle import closed_form as CF
from oracle import functional as OF
from oracle import synth


def test_vocoder_dataflow_matches_oracle(monkeypatch):
    emu_ops.install(monkeypatch)
    g = golden("vocoder_T20.npz")
    m = hsv.Vocoder()
    m.load_state_dict(synth.vocoder_sd(1234), strict=True)
    z, gg = synth.vocoder_inputs(1, 20, seed=1111)
    e, e_ = m.sn(z, gg)
    wav = m(z, gg)
    assert wav.shape == (1, 1, 6400)
    # fp16 operand rounding (emulated) keeps the result inside the parity contract
    assert CF.max_abs(g["wav"], wav.numpy()) <= 2e-3 and CF.snr_db(g["wav"], wav.numpy()) >= 40.0
    assert CF.snr_db(g["e"], e.numpy()) >= 40.0
    assert e_.shape == (1, 1, 80)


def test_speechsr_dataflow_matches_golden(monkeypatch):
    emu_ops.install(monkeypatch)
    for which, cls in ((24, hsv.SpeechSR24), (48, hsv.SpeechSR48)):
        sd = golden_sd(f"speechsr{which}_state.npz")
        g = golden(f"speechsr{which}_example.npz")
        m = cls(100, 40, **hsv.SR_CFG)
        m.load_state_dict(sd, strict=True)
        x = torch.from_numpy(g["x_int16"][:4000].astype(np.float32) / 32768.0).view(1, 1, -1)
        y = m(x).numpy()
        n = y.shape[-1]
        ref = g["y"][..., :n]
        assert CF.max_abs(ref[..., :n - 1500], y[..., :n - 1500]) <= 2e-3
        assert CF.snr_db(ref[..., :n - 1500], y[..., :n - 1500]) >= 40.0


def test_dblock_and_ampblock_dataflow(monkeypatch):
    emu_ops.install(monkeypatch)
    g = golden("dblock_L83.npz")
    sd = synth.hier_generator_sd(1234, "")
    db = hsv.DBlock(64, 512, 4)
    db.load_state_dict({k[len("downs."):]: v for k, v in sd.items() if k.startswith("downs.")}, strict=True)
    y = db(torch.from_numpy(g["x"]))
    assert CF.snr_db(g["y"], y.numpy()) >= 55.0                # gather-then-1x1 == 1x1-then-gather (fp16 operands)
    ga = golden("ampblock_c16_k7.npz")
    gen = torch.Generator().manual_seed(11)
    bsd = {}
    synth._amp_block(bsd, "", gen, 16, 7)
    blk = hsv.AMPBlock1(16, 7, (1, 3, 5))
    blk.load_state_dict(bsd, strict=True)
    x = torch.from_numpy(ga["x"])
    x0 = x.clone()
    y = blk(x)
    assert torch.equal(x, x0)                                  # the input (shared by 3 resblocks) is not modified
    assert CF.snr_db(ga["y"], y.numpy()) >= 50.0


def test_mean_of_blocks_and_cache_invalidation(monkeypatch):
    emu_ops.install(monkeypatch)
    sd = synth.speechsr_sd(7)
    m = hsv.SpeechSR24(100, 40, **hsv.SR_CFG)
    m.load_state_dict(sd, strict=True)
    x = synth.speechsr_input(2, 600)
    y1 = m(x)
    ref1 = OF.speechsr(sd, x, 24)
    assert CF.snr_db(ref1.numpy(), y1.numpy()) >= 40.0
    sd2 = synth.speechsr_sd(8)
    m.load_state_dict(sd2, strict=True)
    y2 = m(x)
    assert CF.snr_db(OF.speechsr(sd2, x, 24).numpy(), y2.numpy()) >= 40.0   # refolded, not stale


def test_front_dataflow_matches_oracle(monkeypatch):
    """The step before the vocoder: the drop-in modules' dataflow (operand packing modes, in-place
    residual updates, fused qkv addressing, flow reversal order, noise draws) on CPU through the op emulation, against
    the oracle restatement of SynthesizerTrn.voice_conversion_noise_control / infer."""
    import megatts2_hierspeechpp_b200 as hsv
    from megatts2_hierspeechpp_b200 import synthetic as synth
    from oracle import functional_front as FF
    emu_ops.install(monkeypatch)
    import megatts2_hierspeechpp_b200.front as FR
    monkeypatch.setattr(FR, "_as_input", lambda x: x.detach().contiguous())
    sd = synth.synthesizer_sd(1234)
    m = hsv.HierSpeechSynthesizer()
    m.load_state_dict(sd, strict=True)
    m.eval()
    T = 16
    w2v, f0, mel = synth.synthesizer_inputs(T, 24, seed=3)
    ln, ln2 = torch.LongTensor([T]), torch.LongTensor([24, 20])           # a padded prompt exercises the masks
    torch.manual_seed(7)
    got = m.voice_conversion_noise_control(w2v, ln, mel, ln2, f0, noise_scale=0.333, denoise_ratio=0.3)
    torch.manual_seed(7)
    ref = FF.voice_conversion_noise_control(sd, w2v, ln, mel, ln2, f0, noise_scale=0.333, denoise_ratio=0.3)
    assert got.shape == ref.shape == (1, 1, 320 * T)
    assert CF.max_abs(ref.numpy(), got.numpy()) <= 2e-3 and CF.snr_db(ref.numpy(), got.numpy()) >= 40.0
    # the front alone, tighter: z before the vocoder
    torch.manual_seed(9)
    z_ref, g_ref = FF.front(sd, w2v, ln, mel, ln2, f0, 0.333, 0.3)
    torch.manual_seed(9)
    tm = torch.unsqueeze(FR.sequence_mask(ln2, mel.size(2)), 1).to(mel.dtype)
    g = m.emb_g(mel, tm).unsqueeze(-1)
    assert (g - FF.style_encoder(sd, "emb_g.", mel, tm).unsqueeze(-1)).abs().max() <= 2e-3


def test_ttv_tail_dataflow_matches_oracle(monkeypatch):
    """W2VDecoder + PitchPredictor (partial) on CPU through the op emulation: split-K conv_pre, the
    leaky_relu packs, the 1/3 of the resblock mean carried into the next pack / the conv_post weight, the accumulating
    resblock epilogues -- against the oracle restatement, with a padded batch."""
    import megatts2_hierspeechpp_b200 as hsv
    from megatts2_hierspeechpp_b200 import synthetic as synth
    from oracle import functional_ttv as FT
    emu_ops.install(monkeypatch)
    import megatts2_hierspeechpp_b200.front as FR
    import megatts2_hierspeechpp_b200.ttv as TV
    for mod in (FR, TV):
        monkeypatch.setattr(mod, "_as_input", lambda x: x.detach().contiguous())
    monkeypatch.setattr(TV.PitchPredictor, "parallel_blocks", False)
    sd = synth.ttv_tail_sd(3456)
    m = hsv.TTVTail()
    m.load_state_dict(sd, strict=True)
    m.eval()
    z, mask, g = synth.ttv_tail_inputs(2, 40, lengths=[40, 31])
    w2v, pitch = m(z, mask, g)
    w2v_ref, pitch_ref = FT.ttv_tail(sd, z, mask, g)
    assert w2v.shape == (2, 1024, 40) and pitch.shape == (2, 1, 160)
    assert CF.snr_db(w2v_ref.numpy(), w2v.numpy()) >= 55.0 and CF.snr_db(pitch_ref.numpy(), pitch.numpy()) >= 55.0
    assert (w2v[1, :, 31:] == 0).all()                                   # masked frames stay exactly zero
    # the pitch predictor alone on the oracle's w2v: isolates it from the decoder's rounding
    p2 = m.pp(w2v_ref, g)
    assert CF.max_abs(pitch_ref.numpy(), p2.numpy()) <= 2e-4
