"""Generate tests/golden/*.npz from the REAL reference (run in the authoring container only).

    python tests/gen_golden.py

Imports the reference's own modules from /root/reference (oracle.refload), runs them on CPU fp32
and stores small input/output vectors.  The reference cannot travel to the GPU box, these fixtures
can.  Real SpeechSR weights are stored as fp32 arrays (they are data, not source).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order
    ref = refload.load()
    H = ref.H

    # ---- SpeechSR real checkpoints + example wav (config #1 known-answer) ----
    wav = refload.example_wav()
    for which in (24, 48):
        m = refload.load_speechsr(which)
        np.savez_compressed(os.path.join(OUT, f"speechsr{which}_state.npz"), **_np(m.state_dict()))
        x = wav if which == 24 else wav[:, :, :16000]
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, f"speechsr{which}_example.npz"),
                            x_int16=np.round(x.numpy().reshape(-1) * 32768.0).astype(np.int16), y=y.numpy())
        print("sr", which, tuple(x.shape), "->", tuple(y.shape), float(y.abs().max()), float(y.sum()))
        if which == 48:
            # the full 3 s example through the 48k twin (fp16 output: 2^-11 relative, far below the 2e-3 bar)
            with torch.no_grad():
                yf = m(wav)
            np.savez_compressed(os.path.join(OUT, "speechsr48_example_full.npz"), y_f16=yf.numpy().astype(np.float16),
                                absmax=np.float32(yf.abs().max()), std=np.float32(yf.std()))
            print("sr 48 full", tuple(yf.shape), float(yf.abs().max()), float(yf.std()))

    # ---- Activation1d cases: odd / even / tiny lengths, checkpoint-like alpha/beta ranges ----
    g = torch.Generator().manual_seed(7)
    cases = {}
    for name, (B, C, L) in dict(even=(2, 8, 64), odd=(1, 5, 37), tiny=(2, 3, 5), one=(1, 2, 1), two=(1, 2, 2),
                                tile=(1, 8, 600)).items():
        act = ref.alias_free_torch.Activation1d(activation=ref.activations.SnakeBeta(C, alpha_logscale=True))
        act.act.alpha.data = torch.rand(C, generator=g) * 3.4 - 1.0    # [-1, 2.4]
        act.act.beta.data = torch.rand(C, generator=g) * 3.7 - 2.9     # [-2.9, 0.8]
        x = torch.randn(B, C, L, generator=g) * 2.0
        with torch.no_grad():
            y = act(x)
        cases[f"{name}_x"] = x.numpy(); cases[f"{name}_y"] = y.numpy()
        cases[f"{name}_alpha"] = act.act.alpha.data.numpy(); cases[f"{name}_beta"] = act.act.beta.data.numpy()
    np.savez_compressed(os.path.join(OUT, "activation1d_cases.npz"), **cases)

    # ---- AMPBlock1 / DBlock with seeded synthetic weights ----
    gen = torch.Generator().manual_seed(11)
    sd = {}
    synth._amp_block(sd, "", gen, 16, 7)
    blk = H.AMPBlock1(16, 7, (1, 3, 5), activation="snakebeta")
    blk.load_state_dict(sd, strict=True); blk.eval()
    x = torch.randn(2, 16, 300, generator=gen)
    with torch.no_grad():
        y = blk(x)
    np.savez_compressed(os.path.join(OUT, "ampblock_c16_k7.npz"), x=x.numpy(), y=y.numpy(), seed=11)

    sd = synth.hier_generator_sd(1234, "")
    dsd = {k[len("downs."):]: v for k, v in sd.items() if k.startswith("downs.")}
    db = H.DBlock(64, 512, 4); db.load_state_dict(dsd, strict=True); db.eval()
    x = torch.randn(1, 64, 83, generator=gen)   # 83 // 4 = 20: non-divisible length
    with torch.no_grad():
        y = db(x)
    np.savez_compressed(os.path.join(OUT, "dblock_L83.npz"), x=x.numpy(), y=y.numpy())

    # ---- vocoder (sn + dec), synthetic weights seed 1234, T=20 frames ----
    vsd = synth.vocoder_sd(1234)
    G = H.Generator(**synth.HIER_CFG); S = H.SourceNetwork(256)
    G.load_state_dict({k[4:]: v for k, v in vsd.items() if k.startswith("dec.")}, strict=True)
    S.load_state_dict({k[3:]: v for k, v in vsd.items() if k.startswith("sn.")}, strict=True)
    G.eval(); S.eval()
    z, gg = synth.vocoder_inputs(1, 20, seed=1111)
    with torch.no_grad():
        e, e_ = S(z, gg)
        o = G(z, e, gg)
    np.savez_compressed(os.path.join(OUT, "vocoder_T20.npz"), e=e.numpy(), e_pred=e_.numpy(), wav=o.numpy())
    print("vocoder", tuple(o.shape), float(o.std()), float(o.abs().max()))

    # ---- interpolation index probes from ATen (CPU): arange input exposes i0/lam ----
    probes = {}
    for Lin, Lout in ((16, 24), (16, 48), (48000, 72000), (160000, 480000), (33, 49), (7, 21)):
        ar = torch.arange(Lin, dtype=torch.float32).view(1, 1, -1)
        v = torch.nn.functional.interpolate(ar, Lout, mode="linear").numpy().reshape(-1)
        probes[f"lin_{Lin}_{Lout}"] = v if Lout <= 4096 else v[::97]   # large cases: every 97th sample
    for Lin, Lout in ((83, 20), (2000, 500), (6000, 1500), (10, 3)):
        ar = torch.arange(Lin, dtype=torch.float32).view(1, 1, -1)
        probes[f"near_{Lin}_{Lout}"] = torch.nn.functional.interpolate(ar, size=Lout).numpy().reshape(-1)
    np.savez_compressed(os.path.join(OUT, "interp_probes.npz"), **probes)

    # ---- state_dict key/shape manifests of the reference modules ----
    man = {}
    for name, mod in (("dec", G), ("sn", S), ("sr", refload.load_speechsr(24))):
        man[name] = np.array([f"{k}:{','.join(map(str, v.shape))}" for k, v in mod.state_dict().items()])
    np.savez_compressed(os.path.join(OUT, "state_dict_manifest.npz"), **man)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
