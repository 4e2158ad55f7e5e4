"""Pin the oracle against the REAL reference modules (authoring container only; skipped on the GPU box)."""
import pytest
import torch

from oracle import functional as OF
from oracle import refload, synth

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def test_vocoder_bit_exact_vs_reference():
    ref = refload.load()
    sd = synth.vocoder_sd(1234)
    G = ref.H.Generator(**synth.HIER_CFG)
    S = ref.H.SourceNetwork(256)
    G.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("dec.")}, strict=True)
    S.load_state_dict({k[3:]: v for k, v in sd.items() if k.startswith("sn.")}, strict=True)
    G.eval(); S.eval()
    z, g = synth.vocoder_inputs(2, 12, seed=5)
    with torch.no_grad():
        e, e_ = S(z, g)
        o = G(z, e, g)
    e2, e2_ = OF.source_network(sd, "sn.", z, g)
    o2 = OF.hier_generator(sd, "dec.", z, e2, g)
    assert torch.equal(e, e2) and torch.equal(e_, e2_) and torch.equal(o, o2)


@pytest.mark.parametrize("which", [24, 48])
def test_speechsr_bit_exact_vs_reference(which):
    m = refload.load_speechsr(which)
    x = refload.example_wav()[:, :, 4000:9000]
    with torch.no_grad():
        a = m(x)
    b = OF.speechsr(m.state_dict(), x, which)
    assert a.shape[-1] == OF.speechsr_out_len(5000, which)
    assert torch.equal(a, b)


def test_synthetic_speechsr_sd_loads_strict():
    ref = refload.load()
    sd = synth.speechsr_sd()
    m = ref.sr24.SynthesizerTrn(100, 40, **synth.SR_CFG)
    m.load_state_dict(sd, strict=True)


def test_b200_modules_are_drop_in_for_reference_state_dicts():
    """Every key/shape of the reference modules' state_dict exists in the B200 modules and vice versa."""
    import megatts2_hierspeechpp_b200 as hsv
    ref = refload.load()
    pairs = [
        (ref.H.Generator(**synth.HIER_CFG), hsv.Generator(**hsv.HIER_CFG)),
        (ref.H.SourceNetwork(256), hsv.SourceNetwork(256)),
        (ref.sr24.SynthesizerTrn(100, 40, **synth.SR_CFG), hsv.SpeechSR24(100, 40, **hsv.SR_CFG)),
        (ref.sr48.SynthesizerTrn(128, 40, **synth.SR_CFG), hsv.SpeechSR48(128, 40, **hsv.SR_CFG)),
        (ref.H.AMPBlock1(32, 7, (1, 3, 5), activation="snakebeta"), hsv.AMPBlock1(32, 7, (1, 3, 5))),
        (ref.H.DBlock(64, 512, 4), hsv.DBlock(64, 512, 4)),
    ]
    for r, m in pairs:
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys()), type(r).__name__
        for k in rs:
            assert rs[k].shape == ms[k].shape, k
        m.load_state_dict(rs, strict=True)
        r.load_state_dict(m.state_dict(), strict=True)


def test_patch_reference_swaps_classes():
    import megatts2_hierspeechpp_b200 as hsv
    ref = refload.load()
    saved = {n: getattr(ref.H, n) for n in ("Generator", "SourceNetwork", "AMPBlock1", "DBlock", "Activation1d",
                                            "PosteriorSFEncoder", "ResidualCouplingBlock_Transformer", "StyleEncoder")}
    saved_sr = {n: getattr(ref.sr24, n) for n in ("Generator", "AMPBlock0", "Activation1d")}
    saved_sr48 = {n: getattr(ref.sr48, n) for n in ("Generator", "AMPBlock0", "Activation1d")}
    try:
        patched = hsv.patch_reference()
        assert "hierspeechpp_speechsynthesizer.Generator" in patched
        assert ref.H.Generator is hsv.Generator and ref.sr24.Generator is hsv.SpeechSR24Generator
        assert ref.H.PosteriorSFEncoder is hsv.front.PosteriorSFEncoder and ref.H.StyleEncoder is hsv.front.StyleEncoder
        m = ref.sr24.SynthesizerTrn(100, 40, **synth.SR_CFG)          # the reference's own SynthesizerTrn
        assert isinstance(m.dec, hsv.SpeechSR24Generator)
        m.load_state_dict(refload.load_speechsr(24).state_dict(), strict=True)
    finally:
        for n, v in saved.items():
            setattr(ref.H, n, v)
        for n, v in saved_sr.items():
            setattr(ref.sr24, n, v)
        for n, v in saved_sr48.items():
            setattr(ref.sr48, n, v)
        import importlib, sys
        for modname in ("alias_free_torch", "activations"):
            importlib.reload(sys.modules[modname])


def test_front_oracle_bit_exact_vs_reference_synthesizer():
    """The step before the vocoder: seeded synthetic front checkpoint loads strictly into the
    reference's SynthesizerTrn (with its other sub-modules untouched), and the oracle restatement of
    voice_conversion_noise_control reproduces the reference bit for bit on CPU."""
    from oracle import functional_front as FF
    ref = refload.load()
    torch.manual_seed(0)
    m = ref.H.SynthesizerTrn(**refload.HIER_SYNTH_CFG)
    sd = synth.synthesizer_sd(1234)
    own = m.state_dict()
    for prefix in ("enc_p_l.", "flow_l.", "flow.", "emb_g.", "sn.", "dec."):
        want = [k for k in own if k.startswith(prefix)]
        got = [k for k in sd if k.startswith(prefix)]
        assert sorted(want) == sorted(got), (prefix, [k for k in want if k not in got][:3], [k for k in got if k not in want][:3])
        for k in want:
            assert own[k].shape == sd[k].shape, k
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.split(".")[0] in ("enc_p", "enc_q", "mel_decoder") for k in missing)
    m.eval()
    T = 24
    w2v, f0, mel = synth.synthesizer_inputs(T, 40, seed=3)
    ln, ln2 = torch.LongTensor([T]), torch.LongTensor([40, 40])
    with torch.no_grad():
        torch.manual_seed(7)
        a = m.voice_conversion_noise_control(w2v, ln, mel, ln2, f0, noise_scale=0.333, denoise_ratio=0.3)
        torch.manual_seed(7)
        b = FF.voice_conversion_noise_control(sd, w2v, ln, mel, ln2, f0, noise_scale=0.333, denoise_ratio=0.3)
    assert a.shape == (1, 1, 320 * T)
    assert torch.equal(a, b)


def test_ttv_tail_oracle_bit_exact_vs_reference_and_drop_in():
    """The tail of the text-to-vec model (partial): the synthetic W2VDecoder / PitchPredictor
    checkpoints load strictly into the reference's classes, the oracle restatement reproduces them bit for bit on CPU,
    the B200 classes have the same state_dict keys and shapes, and patch_reference() swaps them in."""
    import megatts2_hierspeechpp_b200 as hsv
    from oracle import functional_ttv as FT
    T = refload.load_ttv()
    w = T.W2VDecoder(**refload.W2V_DECODER_CFG).eval()
    p = T.PitchPredictor().eval()
    sd = synth.ttv_tail_sd(3456)
    sdw = {k[len("w2v_decoder."):]: v for k, v in sd.items() if k.startswith("w2v_decoder.")}
    sdp = {k[len("pp."):]: v for k, v in sd.items() if k.startswith("pp.")}
    w.load_state_dict(sdw, strict=True)
    p.load_state_dict(sdp, strict=True)
    z, mask, g = synth.ttv_tail_inputs(2, 40, lengths=[40, 31])
    with torch.no_grad():
        a = w(z, mask, g=g)
        b = p(a, g)
    a2, b2 = FT.ttv_tail(sd, z, mask, g)
    assert a.shape == (2, 1024, 40) and b.shape == (2, 1, 160)
    assert torch.equal(a, a2) and torch.equal(b, b2)
    # drop-in: same keys, same order, same shapes
    mine = hsv.TTVTail()
    for r, m in ((w, mine.w2v_decoder), (p, mine.pp)):
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys()), type(r).__name__
        for k in rs:
            assert rs[k].shape == ms[k].shape, k
    mine.load_state_dict(sd, strict=True)
    saved = (T.W2VDecoder, T.PitchPredictor)
    ref = refload.load()
    saved_h = {n: getattr(ref.H, n) for n in ("Generator", "SourceNetwork", "AMPBlock1", "DBlock", "Activation1d",
                                              "PosteriorSFEncoder", "ResidualCouplingBlock_Transformer", "StyleEncoder")}
    saved_sr = {n: getattr(ref.sr24, n) for n in ("Generator", "AMPBlock0", "Activation1d")}
    saved_sr48 = {n: getattr(ref.sr48, n) for n in ("Generator", "AMPBlock0", "Activation1d")}
    try:
        patched = hsv.patch_reference()
        assert "ttv_v1.t2w2v_transformer.PitchPredictor" in patched
        assert T.W2VDecoder is hsv.W2VDecoder and T.PitchPredictor is hsv.PitchPredictor
    finally:
        T.W2VDecoder, T.PitchPredictor = saved
        for n, v in saved_h.items():
            setattr(ref.H, n, v)
        for n, v in saved_sr.items():
            setattr(ref.sr24, n, v)
        for n, v in saved_sr48.items():
            setattr(ref.sr48, n, v)
        import importlib, sys
        for modname in ("alias_free_torch", "activations"):
            importlib.reload(sys.modules[modname])
