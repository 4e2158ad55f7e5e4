"""The tail of the text-to-vec model (partial) on the GPU: ``W2VDecoder`` + ``PitchPredictor`` against the
oracle restatement (oracle/functional_ttv.py, pinned bit-exact to the reference classes on CPU) run in strict fp32 on the
same device, and the chain into ``HierSpeechSynthesizer.voice_conversion_noise_control`` (inference.py:158-167)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import MAX_ABS_TOL, SNR_DB_MIN
from oracle import closed_form as CF
from oracle import functional_front as FF
from oracle import functional_ttv as FT
from oracle import synth
from test_gpu_front import _rel, strict_fp32

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def tail(hsv):
    m = hsv.TTVTail()
    m.load_state_dict(synth.ttv_tail_sd(3456), strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("L", [1, 5, 203, 4096])
def test_direct_conv_input_activations(hsv, L):
    """leaky_relu(0.01) / leaky_relu(0.1) / SiLU on the input of the fp32 direct conv, every kernel variant (row-dot
    for L = 1, thin4 for Cout = 1 k = 7, generic thin, tiled)."""
    ops = hsv.ops
    g = torch.Generator().manual_seed(L)
    for cin, cout, k in ((64, 1, 7), (48, 3, 5), (40, 24, 3)):
        x = torch.randn(2, cin, L, generator=g).to(DEV)
        w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(DEV)
        b = torch.randn(cout, generator=g).to(DEV)
        for flag, fn in ((ops.CONV_LRELU001_IN, lambda t: F.leaky_relu(t, 0.01)),
                         (ops.CONV_LRELU_IN, lambda t: F.leaky_relu(t, 0.1)), (ops.CONV_SILU_IN, F.silu)):
            with strict_fp32():
                ref = F.conv1d(fn(x), w, b, padding=k // 2)
            got = ops.conv1d_direct(x, w, b, pad=k // 2, flags=flag)
            assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), (cin, cout, k, flag)


@pytest.mark.parametrize("B,T,lens", [(1, 200, None), (2, 77, [77, 50]), (4, 333, [333, 300, 129, 1])])
def test_ttv_tail_vs_oracle(hsv, tail, B, T, lens):
    sd = {k: v.to(DEV) for k, v in synth.ttv_tail_sd(3456).items()}
    z, mask, g = (t.to(DEV) for t in synth.ttv_tail_inputs(B, T, seed=11, lengths=lens))
    with strict_fp32(), torch.no_grad():
        w_ref, p_ref = FT.ttv_tail(sd, z, mask, g)
        w, p = tail(z, mask, g)
        p_only = tail.pp(w_ref, g)                       # the pitch predictor on the oracle's input
    assert w.shape == (B, 1024, T) and p.shape == (B, 1, 4 * T)
    snr_w = CF.snr_db(w_ref.cpu().numpy(), w.cpu().numpy())
    snr_p = CF.snr_db(p_ref.cpu().numpy(), p.cpu().numpy())
    print(f"[parity] W2VDecoder B={B} T={T}: rel={_rel(w, w_ref):.2e} snr={snr_w:.1f} dB   PitchPredictor: "
          f"rel={_rel(p, p_ref):.2e} snr={snr_p:.1f} dB  (alone: rel={_rel(p_only, p_ref):.2e})")
    assert _rel(w, w_ref) <= 2e-3 and snr_w >= 55.0
    assert _rel(p, p_ref) <= 2e-3 and snr_p >= 55.0 and _rel(p_only, p_ref) <= 1e-3
    if lens is not None:
        for b, n in enumerate(lens):
            assert bool((w[b, :, n:] == 0).all())        # masked frames are exactly zero, as in the reference
    # one stream per resblock == sequential, bit for bit (accumulation order is chained with events)
    old = type(tail.pp).parallel_blocks
    try:
        type(tail.pp).parallel_blocks = not old
        with torch.no_grad():
            p_seq = tail.pp(w_ref, g)
    finally:
        type(tail.pp).parallel_blocks = old
    assert torch.equal(p_seq, p_only)


def test_ttv_tail_graph_replay_and_chain(hsv, tail):
    """(z, mask, g) -> (w2v, pitch) -> waveform: the two halves replay as CUDA graphs, and the chained waveform meets
    the path's bar against the oracle chain run in strict fp32 (the pitch goes through enc_p_l.pre_filter, a smooth
    map; the log(55 Hz) threshold of inference.py:163 is a caller-side edit and not applied here)."""
    T = 150
    sd_t = {k: v.to(DEV) for k, v in synth.ttv_tail_sd(3456).items()}
    sd_s = {k: v.to(DEV) for k, v in synth.synthesizer_sd(1234).items()}
    syn = hsv.HierSpeechSynthesizer()
    syn.load_state_dict(synth.synthesizer_sd(1234), strict=True)
    syn = syn.to(DEV).eval()
    z, mask, g = (t.to(DEV) for t in synth.ttv_tail_inputs(1, T, seed=5))
    _, _, mel = synth.synthesizer_inputs(T, 150, seed=1111)
    mel = mel.to(DEV)
    ln, ln2 = torch.LongTensor([T]).to(DEV), torch.LongTensor([150, 150]).to(DEV)
    with strict_fp32(), torch.no_grad():
        w_ref, p_ref = FT.ttv_tail(sd_t, z, mask, g)
        torch.manual_seed(7)
        ref = FF.voice_conversion_noise_control(sd_s, w_ref, ln, mel, ln2, p_ref, noise_scale=0.333, denoise_ratio=0.3)
        w, p = tail(z, mask, g)
        torch.manual_seed(7)
        got = syn.voice_conversion_noise_control(w, ln, mel, ln2, p, noise_scale=0.333, denoise_ratio=0.3)
    ma, snr = CF.max_abs(ref.cpu().numpy(), got.cpu().numpy()), CF.snr_db(ref.cpu().numpy(), got.cpu().numpy())
    print(f"[parity] ttv tail -> voice_conversion_noise_control T={T}: max_abs={ma:.3e} snr={snr:.1f} dB")
    assert got.shape == (1, 1, 320 * T) and ma <= MAX_ABS_TOL and snr >= SNR_DB_MIN
    runner = hsv.CudaGraphRunner(tail)
    w_g, p_g = runner(z, mask, g)
    assert torch.equal(w_g, w) and torch.equal(p_g, p)
    w_g2, p_g2 = runner(z, mask, g)                      # second call = pure replay
    assert runner.captures == 1 and torch.equal(p_g2, p)
