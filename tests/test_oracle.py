"""Oracle self-consistency and pinning against the committed golden vectors (CPU)."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_sd
from oracle import closed_form as CF
from oracle import functional as OF
from oracle import synth


def test_filter_taps_match_checkpoints_and_formula():
    sd = golden_sd("speechsr24_state.npz")
    taps = torch.from_numpy(CF.FILTER_TAPS_F32)
    n = 0
    for k, v in sd.items():
        if k.endswith(".filter"):
            assert torch.equal(v.flatten(), taps), k
            n += 1
    assert n == 38  # 19 Activation1d x (up, down)
    assert np.abs(CF.kaiser_sinc_filter1d(0.25, 0.3, 12) - CF.FILTER_TAPS_F32.astype(np.float64)).max() < 5e-8
    assert np.allclose(CF.FILTER_TAPS_F32, CF.FILTER_TAPS_F32[::-1])  # symmetric


@pytest.mark.parametrize("case", ["even", "odd", "tiny", "one", "two", "tile"])
def test_activation1d_golden(case):
    g = golden("activation1d_cases.npz")
    x, y = g[f"{case}_x"], g[f"{case}_y"]
    al, be = g[f"{case}_alpha"], g[f"{case}_beta"]
    sd = {"a.act.alpha": torch.from_numpy(al), "a.act.beta": torch.from_numpy(be),
          "a.upsample.filter": torch.from_numpy(CF.FILTER_TAPS_F32.copy()).view(1, 1, 12),
          "a.downsample.lowpass.filter": torch.from_numpy(CF.FILTER_TAPS_F32.copy()).view(1, 1, 12)}
    yo = OF.activation1d(sd, "a.", torch.from_numpy(x)).numpy()
    assert np.abs(yo - y).max() <= 2e-6            # same ATen ops (thread count may differ)
    yc = CF.activation1d(x, al, be)                # independent fp64 closed form
    assert np.abs(yc - y).max() <= 2e-5 * max(1.0, np.abs(y).max())


def test_ampblock_golden():
    g = golden("ampblock_c16_k7.npz")
    gen = torch.Generator().manual_seed(11)
    sd = {}
    synth._amp_block(sd, "", gen, 16, 7)
    y = OF.amp_block(sd, "", torch.from_numpy(g["x"]), 7).numpy()
    assert np.abs(y - g["y"]).max() <= 1e-5


def test_dblock_golden():
    g = golden("dblock_L83.npz")
    sd = synth.hier_generator_sd(1234, "")
    y = OF.dblock(sd, "downs.", torch.from_numpy(g["x"])).numpy()
    assert y.shape == g["y"].shape == (1, 512, 20)
    assert np.abs(y - g["y"]).max() <= 1e-5


def test_vocoder_golden():
    g = golden("vocoder_T20.npz")
    sd = synth.vocoder_sd(1234)
    z, gg = synth.vocoder_inputs(1, 20, seed=1111)
    e, e_ = OF.source_network(sd, "sn.", z, gg)
    wav = OF.hier_generator(sd, "dec.", z, e, gg)
    assert wav.shape == (1, 1, 6400)
    assert np.abs(e.numpy() - g["e"]).max() <= 1e-5
    assert np.abs(wav.numpy() - g["wav"]).max() <= 1e-5


@pytest.mark.parametrize("which", [24, 48])
def test_speechsr_golden(which):
    sd = golden_sd(f"speechsr{which}_state.npz")
    g = golden(f"speechsr{which}_example.npz")
    x = torch.from_numpy(g["x_int16"].astype(np.float32) / 32768.0).view(1, 1, -1)
    if which == 24:
        x = x[:, :, :8000]          # keep the CPU suite short; the full 3 s runs on the GPU
    y = OF.speechsr(sd, x, which).numpy()
    n = y.shape[-1]
    # interior samples are independent of the truncation (receptive field << 1000 samples)
    assert np.abs(y[..., : n - 1000] - g["y"][..., : n - 1000]).max() <= 1e-5


def test_interp_tables_golden():
    g = golden("interp_probes.npz")
    for key in g.files:
        kind, lin, lout = key.split("_")
        lin, lout = int(lin), int(lout)
        if kind == "lin":
            i0, i1, lam = CF.linear_interp_table(lin, lout, fma=False)
            v = ((1.0 - lam) * i0.astype(np.float32) + lam * i1.astype(np.float32)).astype(np.float32)
            ref = g[key]
            if lout > 4096:
                v = v[::97]
            assert np.abs(v - ref).max() <= 1e-3 * max(1.0, lin / 1e4), key
            assert (i1 - i0).max() <= 1 and i1.max() == lin - 1
        else:
            assert np.array_equal(CF.nearest_index(lin, lout).astype(np.float32), g[key]), key


def test_index_tables_integer_properties():
    for L in (1, 2, 5, 37, 64):
        ui, di = CF.up_indices(L), CF.down_indices(L)
        assert ui.min() == 0 and ui.max() == L - 1 and di.min() == 0 and di.max() == 2 * L - 1
        if L > 12:
            assert np.array_equal(di[6], np.arange(12) + 7)       # interior: z[2t-5 .. 2t+6]
            assert np.array_equal(ui[2 * 6], 6 + 2 - np.arange(6))  # interior even sample
    for k, u in ((8, 4), (11, 5), (4, 2)):
        table = CF.conv_transpose_phase_table(k, u)
        taps = sorted(j for ph in table for j, _ in ph)
        assert taps == list(range(k))                           # every tap used by exactly one phase
        assert max(len(ph) for ph in table) <= 3


def test_conv_closed_forms_match_aten():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 40, generator=g)
    w = torch.randn(5, 6, 7, generator=g) * 0.1
    b = torch.randn(5, generator=g)
    for d in (1, 3, 5):
        ref = torch.nn.functional.conv1d(x, w, b, padding=OF.get_padding(7, d), dilation=d).numpy()
        assert np.abs(CF.conv1d(x.numpy(), w.numpy(), b.numpy(), d) - ref).max() < 1e-5
    for k, u in ((8, 4), (11, 5), (4, 2)):
        wt = torch.randn(6, 3, k, generator=g) * 0.1
        bt = torch.randn(3, generator=g)
        ref = torch.nn.functional.conv_transpose1d(x, wt, bt, stride=u, padding=(k - u) // 2).numpy()
        out = CF.conv_transpose1d(x.numpy(), wt.numpy(), bt.numpy(), u)
        assert out.shape == ref.shape == (2, 3, 40 * u)
        assert np.abs(out - ref).max() < 1e-5


def test_snr_metric():
    a = np.sin(np.arange(1000) * 0.01)
    assert CF.snr_db(a, a) == float("inf")
    assert abs(CF.snr_db(a, a * 1.01) - 40.0) < 0.1


def test_oracle_peak_norm_pcm16_matches_reference_expression():
    """oracle.functional.peak_norm_pcm16 is the literal expression of inference_plm.py:183-188 /
    inference_speechsr.py:39-41; check order sensitivity is preserved (the two scripts differ in operand order)."""
    import numpy as np
    import torch
    from oracle import functional as OF
    g = torch.Generator().manual_seed(0)
    a = torch.tanh(torch.randn(1, 1, 50000, generator=g))
    plm = (a.squeeze() / torch.abs(a.squeeze()).max() * 32767.0 * 0.999).numpy().astype("int16")
    sr = (a.squeeze() / torch.abs(a.squeeze()).max() * 0.999 * 32767.0).numpy().astype("int16")
    assert np.array_equal(OF.peak_norm_pcm16(a, 32767.0, 0.999), plm)
    assert np.array_equal(OF.peak_norm_pcm16(a, 0.999, 32767.0), sr)
    assert np.abs(plm.astype(np.int32) - sr.astype(np.int32)).max() <= 1
