#!/usr/bin/env python
"""Benchmark of the waveform-generation hot path (BASELINE.json metric: vocoder audio-sec/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload vocoder|speechsr48|speechsr24] [--batch B] [--seconds S]

Default workload = BASELINE.json configs[1]: HierSpeech++ SourceNetwork + Generator (libritts960 arch,
seeded random init), B=1 x 10 s (z [1,192,500], g [1,256,1]) -> 16 kHz wav [1,1,160000], one GPU.
A "step" is one pass of the hot path over one batch.  N>1 (torchrun, one rank per GPU): every rank
runs its own utterances (weak scaling, no data-path collective); NCCL is used for the barrier and the
max-over-ranks of the device time only.

`--impl reference` times the reference's CPU implementation of the same step (the oracle port: the
reference's own ATen op sequence, fp32, all host threads) and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0
TENSOR_FALLBACK_TFLOPS = 1590.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return HBM_FALLBACK_GBS, TENSOR_FALLBACK_TFLOPS, "fallback"


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def make_workload(args, rank):
    """Returns dict(name, sd, host_inputs(list of CPU tensors), audio_seconds, which)."""
    from megatts2_hierspeechpp_b200 import synthetic as synth
    import numpy as np

    if args.workload == "vocoder":
        T = int(round(args.seconds * 50))
        z, g = synth.vocoder_inputs(args.batch, T, seed=1111 + rank)
        return dict(name=f"hierspeechpp_vocoder_sn+dec_B{args.batch}x{args.seconds:g}s", kind="vocoder",
                    sd=synth.vocoder_sd(1234), host_inputs=[z, g], audio_seconds=args.batch * T / 50.0,
                    data="synthetic z~N(0,1) [B,192,T], g~N(0,1) [B,256,1] (seed 1111+rank); "
                         "random-init weights (seed 1234, SnakeBeta alpha~U(-0.5,1), beta~U(-0.5,0.8))")
    if args.workload == "chain24":
        # config #4's timed stage: vocoder (random init) -> SpeechSR24 (bundled checkpoint), one CUDA graph
        T = int(round(args.seconds * 50))
        z, g = synth.vocoder_inputs(args.batch, T, seed=1111 + rank)
        sd = {"vocoder." + k: v for k, v in synth.vocoder_sd(1234).items()}
        gpath = os.path.join(ROOT, "tests", "golden", "speechsr24_state.npz")
        sd.update({"sr." + k: torch.from_numpy(v.copy()) for k, v in np.load(gpath).items()})
        return dict(name=f"config4_vocoder->speechsr24_B{args.batch}x{args.seconds:g}s", kind="chain24", sd=sd,
                    host_inputs=[z, g], audio_seconds=args.batch * T / 50.0, which=24,
                    data="synthetic z, g as config #2; vocoder random init (seed 1234) + bundled speechsr24k checkpoint")
    if args.workload == "synth":
        # config #2 at the SynthesizerTrn level: w2v + f0 + prompt mel -> 16 kHz wav through
        # voice_conversion_noise_control: StyleEncoder, PosteriorSFEncoder, both reverse flows, sn, dec
        if args.batch != 1:
            raise SystemExit("the synth workload is the reference's single-utterance call (B = 1)")
        T = int(round(args.seconds * 50))
        w2v, f0, mel = synth.synthesizer_inputs(T, 150, seed=1111 + rank)
        return dict(name=f"hierspeechpp_synthesizer_vc_noise_control_B1x{args.seconds:g}s", kind="synth",
                    sd=synth.synthesizer_sd(1234), host_inputs=[w2v, f0, mel], audio_seconds=T / 50.0, T=T,
                    data="synthetic w2v~N(0,1) [1,1024,T], f0=log(hz+1) 30 % unvoiced [1,1,4T], prompt mel~N(-4,2) [2,80,150]; "
                         "random-init weights (seed 1234; flows' post / adaLN small random instead of zero)")
    if args.workload == "tts":
        # the device-resident tail of text-to-speech (inference.py:158-167): the text-to-vec model's flow output z ->
        # W2VDecoder -> PitchPredictor -> voice_conversion_noise_control -> 16 kHz wav (SURVEY.md §8f4 partial + f2 + a)
        if args.batch != 1:
            raise SystemExit("the tts workload is the reference's single-utterance call (B = 1)")
        T = int(round(args.seconds * 50))
        z, mask, g = synth.ttv_tail_inputs(1, T, seed=1111 + rank)
        _, _, mel = synth.synthesizer_inputs(T, 150, seed=1111 + rank)
        sd = {"syn." + k: v for k, v in synth.synthesizer_sd(1234).items()}
        sd.update({"tail." + k: v for k, v in synth.ttv_tail_sd(3456).items()})
        return dict(name=f"ttv_tail->synthesizer_vc_noise_control_B1x{args.seconds:g}s", kind="tts", sd=sd,
                    host_inputs=[z, g, mel], audio_seconds=T / 50.0, T=T,
                    data="synthetic z~N(0,1) [1,256,T], g~0.5*N(0,1) [1,256,1], prompt mel~N(-4,2) [2,80,150]; random-init "
                         "weights (seeds 1234 / 3456)")
    which = 48 if args.workload == "speechsr48" else 24
    L = int(round(args.seconds * 16000))
    x = synth.speechsr_input(args.batch, L, seed=1111 + rank)
    gpath = os.path.join(ROOT, "tests", "golden", f"speechsr{which}_state.npz")
    sd = {k: torch.from_numpy(v.copy()) for k, v in np.load(gpath).items()}
    return dict(name=f"speechsr{which}_B{args.batch}x{args.seconds:g}s", kind=f"sr{which}", sd=sd, host_inputs=[x],
                audio_seconds=args.batch * L / 16000.0, which=which,
                data=f"synthetic 0.1*N(0,1) [B,1,L] (seed 1111+rank); bundled speechsr{which}k checkpoint weights")


def oracle_forward(wl, device="cpu"):
    """The oracle port's forward of a workload (the reference's ATen op sequence, oracle/functional.py)."""
    from oracle import functional as OF

    sd = {k: v.to(device) for k, v in wl["sd"].items()}
    ins = [t.to(device) for t in wl["host_inputs"]]
    if wl["kind"] == "vocoder":
        return lambda: OF.vocoder(sd, ins[0], ins[1])
    if wl["kind"] == "synth":
        from oracle import functional_front as FF
        T = wl["T"]
        ln, ln2 = torch.LongTensor([T]).to(device), torch.LongTensor([150, 150]).to(device)
        return lambda: FF.voice_conversion_noise_control(sd, ins[0], ln, ins[2], ln2, ins[1], 0.333, 0.3)
    if wl["kind"] == "tts":
        from oracle import functional_front as FF
        from oracle import functional_ttv as FT
        T = wl["T"]
        st = {k[len("tail."):]: v for k, v in sd.items() if k.startswith("tail.")}
        sy = {k[len("syn."):]: v for k, v in sd.items() if k.startswith("syn.")}
        ln, ln2 = torch.LongTensor([T]).to(device), torch.LongTensor([150, 150]).to(device)
        mask = torch.ones(1, 1, T, device=device)

        def fwd():
            w2v, pitch = FT.ttv_tail(st, ins[0], mask, ins[1])
            return FF.voice_conversion_noise_control(sy, w2v, ln, ins[2], ln2, pitch, 0.333, 0.3)
        return fwd
    if wl["kind"] == "chain24":
        sv = {k[len("vocoder."):]: v for k, v in sd.items() if k.startswith("vocoder.")}
        ss = {k[len("sr."):]: v for k, v in sd.items() if k.startswith("sr.")}
        return lambda: OF.speechsr(ss, OF.vocoder(sv, ins[0], ins[1]), 24)
    return lambda: OF.speechsr(sd, ins[0], wl["which"])


def reference_forward(wl):
    """(callable, kind): the reference's OWN modules when its sources are present (/root/reference here,
    baseline/_ref on the GPU box: oracle/refload.py), else the oracle port."""
    try:
        from oracle import refload
        if refload.available():
            ref = refload.load()
            if wl["kind"] == "chain24":
                raise RuntimeError("chain workload: oracle port (the reference chains the two models in a script)")
            if wl["kind"] == "tts":
                TT = refload.load_ttv()
                m = ref.H.SynthesizerTrn(**refload.HIER_SYNTH_CFG)
                m.load_state_dict({k[4:]: v for k, v in wl["sd"].items() if k.startswith("syn.")}, strict=False)
                wd, pp = TT.W2VDecoder(**refload.W2V_DECODER_CFG), TT.PitchPredictor()
                wd.load_state_dict({k[len("tail.w2v_decoder."):]: v for k, v in wl["sd"].items()
                                    if k.startswith("tail.w2v_decoder.")}, strict=True)
                pp.load_state_dict({k[len("tail.pp."):]: v for k, v in wl["sd"].items() if k.startswith("tail.pp.")},
                                   strict=True)
                m.eval(); wd.eval(); pp.eval()
                z, g, mel = wl["host_inputs"]
                ln, ln2 = torch.LongTensor([wl["T"]]), torch.LongTensor([150, 150])
                mask = torch.ones(1, 1, wl["T"])

                def fwd():
                    w2v = wd(z, mask, g=g)
                    return m.voice_conversion_noise_control(w2v, ln, mel, ln2, pp(w2v, g), 0.333, False, 0.3)
                return fwd, "reference"
            if wl["kind"] == "synth":
                m = ref.H.SynthesizerTrn(**refload.HIER_SYNTH_CFG)
                m.load_state_dict(wl["sd"], strict=False)
                m.eval()
                w2v, f0, mel = wl["host_inputs"]
                ln, ln2 = torch.LongTensor([wl["T"]]), torch.LongTensor([150, 150])
                return (lambda: m.voice_conversion_noise_control(w2v, ln, mel, ln2, f0, 0.333, False, 0.3)), "reference"
            if wl["kind"] == "vocoder":
                from megatts2_hierspeechpp_b200.config import HIER_CFG
                G = ref.H.Generator(**HIER_CFG)
                S = ref.H.SourceNetwork(HIER_CFG["upsample_initial_channel"] // 2)
                G.load_state_dict({k[4:]: v for k, v in wl["sd"].items() if k.startswith("dec.")}, strict=True)
                S.load_state_dict({k[3:]: v for k, v in wl["sd"].items() if k.startswith("sn.")}, strict=True)
                G.eval(); S.eval()
                z, g = wl["host_inputs"]

                def fwd():
                    e, _ = S(z, g)
                    return G(z, e, g)
                return fwd, "reference"
            m = refload.load_speechsr(wl["which"])
            x = wl["host_inputs"][0]
            return (lambda: m(x)), "reference"
    except Exception as e:  # a broken copy must not take the arm down: fall back to the port and say so
        print(f"bench.py: reference modules unavailable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
    return oracle_forward(wl), "port"


def time_cpu(fn, steps, warmup):
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(steps):
            t = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t)
    return ts


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm (CPU)
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args, 0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = f"{wl['name']} (the full step)"
    if wl["audio_seconds"] > 40:          # bound each CPU step (cost is linear in batch x duration)
        a2 = argparse.Namespace(**vars(args))
        a2.batch, a2.seconds = 1, min(args.seconds, 10.0)
        wl = make_workload(a2, 0)
        sample = f"{wl['name']} (slice of the workload; CPU cost is linear in batch x duration)"
    fwd, kind = reference_forward(wl)
    ts = time_cpu(fwd, args.steps, args.warmup)
    ms = 1e3 * sum(ts) / len(ts)
    val = wl["audio_seconds"] / (ms / 1e3)
    sample += f", {args.steps} timed steps after {args.warmup} warm-up, mean"
    how = ("the reference's own nn.Modules (baseline/_ref), unmodified, torch CPU fp32" if kind == "reference" else
           "oracle port (same ATen op sequence as the reference modules; reference sources not present)")
    line = {
        "impl": "reference", "metric": "vocoder audio-sec/sec (RTF^-1)", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": wl["data"],
        "config": {"workload": wl["name"], "device": "cpu", "threads": cores, "note": "reference CPU path = " + how},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def build_model(wl, device):
    import megatts2_hierspeechpp_b200 as hsv

    if wl["kind"] == "vocoder":
        m = hsv.Vocoder()
    elif wl["kind"] == "chain24":
        m = hsv.VocoderSR(24)
    elif wl["kind"] == "synth":
        class SynthStep(torch.nn.Module):
            """voice_conversion_noise_control(w2v, [T], mel, [150, 150], f0, 0.333, denoise_ratio 0.3) as a tensor->tensor
            callable (lengths are host constants, as in inference_plm.py:166-173)."""

            def __init__(self, T):
                super().__init__()
                self.net = hsv.HierSpeechSynthesizer()
                self.T = T

            def load_state_dict(self, sd, strict=True):
                return self.net.load_state_dict(sd, strict=strict)

            def forward(self, w2v, f0, mel):
                ln = torch.full((1,), self.T, dtype=torch.long, device=w2v.device)
                ln2 = torch.full((2,), mel.shape[-1], dtype=torch.long, device=w2v.device)
                return self.net.voice_conversion_noise_control(w2v, ln, mel, ln2, f0, 0.333, False, 0.3)
        m = SynthStep(wl["T"])
    elif wl["kind"] == "tts":
        class TTSStep(torch.nn.Module):
            """(z, g, prompt mel) -> wav: TTVTail -> voice_conversion_noise_control, lengths are host constants."""

            def __init__(self, T):
                super().__init__()
                self.tail = hsv.TTVTail()
                self.syn = hsv.HierSpeechSynthesizer()
                self.T = T

            def forward(self, z, g, mel):
                mask = torch.ones(1, 1, self.T, dtype=torch.float32, device=z.device)
                w2v, pitch = self.tail(z, mask, g)
                ln = torch.full((1,), self.T, dtype=torch.long, device=z.device)
                ln2 = torch.full((2,), mel.shape[-1], dtype=torch.long, device=z.device)
                return self.syn.voice_conversion_noise_control(w2v, ln, mel, ln2, pitch, 0.333, False, 0.3)
        m = TTSStep(wl["T"])
    else:
        m = (hsv.SpeechSR48 if wl["which"] == 48 else hsv.SpeechSR24)(100, 40, **hsv.SR_CFG)
    m.load_state_dict(wl["sd"], strict=True)
    m._bench_workload = wl["name"]
    return m.to(device).eval()


FAMILIES = (
    # (family, substrings of the kernel names CUPTI / ncu report, host ops that launch it)
    ("conv_umma", ("conv_umma",), ("conv1d_umma", "conv_transpose1d_umma", "act_conv1d_umma")),
    ("act1d", ("act1d_kernel", "act1d_mma_kernel"), ("act1d", "act1d_blk16")),
    ("pack_blk16", ("pack_blk16_kernel",), ("pack_blk16",)),
    ("frame_ops", ("pack_act_kernel", "ln_mod_kernel", "frame_op_kernel", "mha_kernel", "conv1d_c1_strided",
                   "masked_mean_kernel", "mean_div_kernel", "absmax_kernel", "pcm16_kernel", "rowmax_kernel"),
     ("pack_blk16_act", "ln_mod_blk16", "frame_op", "mha", "conv1d_c1_strided", "masked_mean")),
    ("small_fp32", ("conv1d_thin", "conv1d_rowdot", "conv1d_tiled", "conv_transpose1d_kernel", "add3_bcast",
                    "nearest_gather", "sr_pre_interp", "weight_norm_fold", "pack_weight"),
     ("conv1d_direct", "conv_transpose1d", "add3_bcast", "nearest_gather", "sr_pre_interp")),
)


def family_of_kernel(name):
    for fam, subs, _ in FAMILIES:
        if any(sub in name for sub in subs):
            return fam
    return None


def algorithmic_work(model, dev_inputs):
    """One eager forward with every hsv op wrapped: algorithmic bytes and FLOPs per kernel family (DESIGN.md §4:
    activation = read x once + write the result once; conv = fp16 operand + residual + output (+ weights); FLOPs =
    2*B*Cin*Cout*taps*L), and the eager per-launch CUDA-event times (one stream, L2 warm) for reference."""
    from megatts2_hierspeechpp_b200 import ops

    records, originals = [], {}

    def wrap(name, bytes_fn=None, flops_fn=None):
        fn = getattr(ops, name)
        originals[name] = fn

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            records.append((name, e0, e1, bytes_fn(*a, **k) if bytes_fn else 0.0, flops_fn(*a, **k) if flops_fn else 0.0))
            return out

        setattr(ops, name, timed)

    wrap("act1d", bytes_fn=lambda x, *a, **k: 8.0 * x.numel())                 # fp32 in + fp32 out
    wrap("act1d_blk16", bytes_fn=lambda x, *a, **k: 6.0 * x.numel())           # fp32 in + fp16 out
    wrap("pack_blk16", bytes_fn=lambda x, *a, **k: (4.0 * len(x) + 2.0) * x[0].numel() if isinstance(x, (list, tuple))
         else 6.0 * x.numel())                                                 # fp32 addend(s) in + fp16 out

    def umma_bytes(a_blk, w, bias, L, cin, cout, k, d, n_tile, residual=None, out=None, acc=None, acc_mode=0, **kw):
        B = a_blk.shape[0]
        b = 2.0 * B * cin * L + 2.0 * cout * cin * k           # fp16 operand + packed weights
        if residual is not None:
            b += 4.0 * B * cout * L
        if out is not None or kw.get("want_out", True):
            b += 4.0 * B * cout * L
        if acc_mode == 1:
            b += 4.0 * B * cout * L
        elif acc_mode in (2, 3):
            b += 8.0 * B * cout * L
        return b

    wrap("conv1d_umma", bytes_fn=umma_bytes,
         flops_fn=lambda a_blk, w, bias, L, cin, cout, k, d, n_tile, **kw: 2.0 * a_blk.shape[0] * cin * cout * k * L)
    wrap("conv_transpose1d_umma",
         bytes_fn=lambda a_blk, w, bias, Lin, cin, cout, k, u, n_tile, add=None: (
             2.0 * a_blk.shape[0] * cin * Lin + 2.0 * cin * cout * k + 4.0 * a_blk.shape[0] * cout * u * Lin *
             (2 if add is not None else 1)),
         flops_fn=lambda a_blk, w, bias, Lin, cin, cout, k, u, n_tile, add=None: 2.0 * a_blk.shape[0] * cin * cout * k * Lin)
    wrap("conv1d_direct",
         bytes_fn=lambda x, w, *a, **k: 4.0 * (x.numel() + x.shape[0] * w.shape[0] * x.shape[2]),
         flops_fn=lambda x, w, *a, **k: 2.0 * x.shape[0] * w.shape[0] * w.shape[1] * w.shape[2] * x.shape[2])
    wrap("conv_transpose1d",
         flops_fn=lambda x, w, *a, **k: 2.0 * x.shape[0] * w.shape[0] * w.shape[1] * w.shape[2] * x.shape[2])
    wrap("sr_pre_interp", bytes_fn=lambda x, w, b, Lout: 4.0 * (x.numel() + x.shape[0] * w.shape[0] * Lout))
    wrap("nearest_gather", bytes_fn=lambda x, Lout: 8.0 * x.shape[0] * x.shape[1] * Lout)
    wrap("add3_bcast", bytes_fn=lambda a, b, bc, out=None: 4.0 * a.numel() * (3 if b is not None else 2))
    try:
        with torch.no_grad():
            model(*dev_inputs)          # pass 1 warms the eager allocator pool; discarded
            torch.cuda.synchronize()
            records.clear()
            model(*dev_inputs)
        torch.cuda.synchronize()
    finally:
        for n, fn in originals.items():
            setattr(ops, n, fn)
    op2fam = {op: fam for fam, _, opsn in FAMILIES for op in opsn}
    fams = {}
    for name, e0, e1, nbytes, flops in records:
        f = fams.setdefault(op2fam.get(name, "small_fp32"), {"launches": 0, "eager_ms": 0.0, "bytes": 0.0, "flops": 0.0})
        f["launches"] += 1
        f["eager_ms"] += e0.elapsed_time(e1)
        f["bytes"] += nbytes
        f["flops"] += flops
    return fams


def cupti_step_profile(step_fn, flush, reps=3):
    """Per-kernel device times of the step AS BENCHMARKED (CUDA-graph replay, multi-stream, PDL, one L2 flush per
    step), from CUPTI through torch.profiler (kineto).  Returns a list (one entry per replay) of lists of
    (kernel name, start_us, duration_us) restricted to hsv kernels, or None when CUPTI is unavailable."""
    try:
        from torch.profiler import ProfilerActivity, profile

        marks = []
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                flush()
                torch.cuda.synchronize()
                t0 = time.perf_counter_ns()
                step_fn()
                torch.cuda.synchronize()
                marks.append((t0, time.perf_counter_ns()))
        evs = []
        kin = getattr(getattr(prof, "profiler", None), "kineto_results", None)
        if kin is not None:
            for e in kin.events():
                try:
                    if "cuda" not in str(e.device_type()).lower():
                        continue
                    evs.append((e.name(), e.start_ns() / 1e3, e.duration_ns() / 1e3))
                except Exception:
                    continue
        if not evs:
            for e in prof.events():
                if "cuda" in str(getattr(e, "device_type", "")).lower():
                    evs.append((e.name, float(e.time_range.start), float(e.time_range.end - e.time_range.start)))
        if os.environ.get("BENCH_DUMP_KERNELS"):          # debugging aid: every device kernel of the profiled replays
            agg = {}
            for n, st, du in evs:
                a = agg.setdefault(n[:90], [0, 0.0])
                a[0] += 1; a[1] += du
            with open(os.environ["BENCH_DUMP_KERNELS"], "w") as f:
                for n, (c, du) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    f.write(f"{du / reps:10.1f} us/step  n={c / reps:7.1f}  avg={du / c:7.2f} us  {n}\n")
        evs = sorted((n, st, du) for n, st, du in evs if family_of_kernel(n) is not None)
        if not evs:
            return None
        # split into replays at the largest gaps (the synchronise + flush between replays)
        evs.sort(key=lambda t: t[1])
        gaps = sorted(range(1, len(evs)), key=lambda i: -(evs[i][1] - (evs[i - 1][1] + evs[i - 1][2])))[:reps - 1]
        cuts = [0] + sorted(gaps) + [len(evs)]
        return [evs[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    except Exception as e:
        print(f"bench.py: CUPTI step profile unavailable ({type(e).__name__}: {e})", file=sys.stderr)
        return None


def _union_ms(intervals):
    tot, end = 0.0, -1e30
    for st, du in sorted(intervals):
        if st > end:
            tot += du
            end = st + du
        elif st + du > end:
            tot += st + du - end
            end = st + du
    return tot / 1e3


def kernel_roofline(fams, replays, ms_per_step, hbm_peak, tensor_peak, peak_kind, wl_key):
    """`roofline` blocks per kernel family, in the regime that is benchmarked: time = the family's kernel durations
    inside one graph replay (CUPTI; median over the profiled replays), bytes / FLOPs = the algorithmic work of the
    same launches.  Concurrent kernels (three resblock streams) share the GPU, so `busy_ms` (union of the family's
    intervals) <= ms_per_step while `sum_ms` may exceed it; `achieved` uses `sum_ms` (per-kernel view) and
    `achieved_busy` the union (what the family sustains while any of its kernels runs)."""
    traffic, traffic_meta = {}, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic_meta = tj.get("_meta")
        traffic = tj.get(wl_key, {})
    timing = {}
    if replays:
        for fam in fams:
            per = []
            for rep in replays:
                iv = [(st, du) for n, st, du in rep if family_of_kernel(n) == fam]
                per.append((sum(du for _, du in iv) / 1e3, _union_ms(iv), len(iv)))
            per.sort()
            timing[fam] = per[len(per) // 2]
        tot = []
        for rep in replays:
            iv = [(st, du) for _, st, du in rep]
            tot.append((sum(du for _, du in iv) / 1e3, _union_ms(iv), len(iv)))
        tot.sort()
        timing["_all"] = tot[len(tot) // 2]
    entries = {}
    all_sum = timing.get("_all", (sum(f["eager_ms"] for f in fams.values()), None, None))[0]
    for fam, f in fams.items():
        if fam not in ("conv_umma", "act1d"):
            continue
        sum_ms, busy_ms, n_graph = timing.get(fam, (f["eager_ms"], None, None))
        ach = f["bytes"] / (sum_ms * 1e-3) / 1e9 if sum_ms else 0.0
        t = traffic.get("conv1d_umma" if fam == "conv_umma" else "act1d")
        dram = t["dram_bytes_per_launch"] if t else None
        alg_per = f["bytes"] / max(1, f["launches"])
        if dram is not None and dram < 0.8 * alg_per:
            bound, why = "latency", ("measured DRAM traffic per launch is below the algorithmic bytes: at this size the "
                                     "tensors are L2-resident and every launch is < 1 wave; the HBM peak is the stated "
                                     "denominator, not the limiter")
        else:
            bound, why = "hbm", "DRAM traffic ~ algorithmic bytes"
        e = {"kernel": ("conv_umma_kernel (tcgen05/TMEM implicit-GEMM Conv1d / ConvTranspose1d, fp16 operands, fp32 "
                        "accumulate)" if fam == "conv_umma" else
                        "act1d_kernel (fused Activation1d/SnakeBeta, fp32 in, fp32|fp16 out)"),
             "bound": bound, "bound_note": why, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
             "traffic": dram, "traffic_source": traffic_meta,
             "algorithmic_bytes_per_launch": alg_per, "peak_kind": peak_kind, "launches_per_step": f["launches"],
             "timing": "CUPTI kernel durations inside the benchmarked CUDA-graph replay (median of the profiled replays)"
                       if replays else "eager CUDA events (CUPTI unavailable)",
             "sum_ms": sum_ms, "busy_ms": busy_ms, "avg_launch_us": 1e3 * sum_ms / max(1, f["launches"]),
             "achieved_busy": (f["bytes"] / (busy_ms * 1e-3) / 1e9) if busy_ms else None,
             "share_of_step_kernel_time": sum_ms / all_sum if all_sum else None,
             "eager_avg_launch_us": 1e3 * f["eager_ms"] / max(1, f["launches"])}
        if fam == "conv_umma":
            tf = f["flops"] / (sum_ms * 1e-3) / 1e12 if sum_ms else 0.0
            e["tensor"] = {"achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tf / tensor_peak}
        entries[fam] = e
    dom = max(entries, key=lambda k2: entries[k2]["share_of_step_kernel_time"] or 0.0)
    roof = entries.pop(dom)
    shares = {}
    for fam, f in fams.items():
        sum_ms, busy_ms, n_graph = timing.get(fam, (f["eager_ms"], None, None))
        shares[fam] = {"launches": f["launches"], "sum_ms": round(sum_ms, 4),
                       "busy_ms": None if busy_ms is None else round(busy_ms, 4),
                       "share": round(sum_ms / all_sum, 4) if all_sum else None}
    tot_bytes = sum(f["bytes"] for f in fams.values())
    tot_flops = sum(f["flops"] for f in fams.values())
    step = {"algorithmic_bytes": tot_bytes, "flops": tot_flops, "ms_per_step": ms_per_step,
            "hbm": {"achieved": tot_bytes / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": tot_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
            "tensor": {"achieved": tot_flops / (ms_per_step * 1e-3) / 1e12, "peak": tensor_peak, "unit": "TFLOP/s",
                       "frac": tot_flops / (ms_per_step * 1e-3) / 1e12 / tensor_peak},
            "note": "whole step: algorithmic bytes of THIS dataflow (fp16 operand written by the activation and re-read "
                    "by the conv included) and conv FLOPs over the device-timed ms_per_step"}
    if "_all" in timing:
        step["kernel_sum_ms"], step["gpu_busy_ms"], step["kernels_in_graph"] = timing["_all"]
        step["concurrency"] = timing["_all"][0] / timing["_all"][1] if timing["_all"][1] else None
    return roof, entries, shares, step


def saturated_rooflines(dev, hbm_peak, tensor_peak):
    """The two hot kernels timed alone at a size that fills the GPU and exceeds L2 (config #3's per-layer
    shape, batch 16: [16,32,480000]): what the kernels sustain when the workload is not latency-bound."""
    import megatts2_hierspeechpp_b200 as hsv

    B, C, L, k, d = 16, 32, 480000, 7, 3
    x = torch.randn(B, C, L, device=dev)
    a = torch.zeros(C, device=dev)
    buf = hsv.ops.blk16_buffer(B, C, L, dev, slot=5)

    def timeit(fn, n=5):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_act = timeit(lambda: hsv.ops.act1d_blk16(x, a, a, buf))
    w = torch.randn(C, C, k, device=dev) * 0.05
    wp = hsv.ops.pack_conv_weight(w, C)
    out = torch.empty_like(x)
    ms_conv = timeit(lambda: hsv.ops.conv1d_umma(buf, wp, a, L, C, C, k, d, C, residual=x, out=out))
    n = B * C * L
    act_gbs = 6.0 * n / (ms_act * 1e-3) / 1e9
    conv_gbs = 10.0 * n / (ms_conv * 1e-3) / 1e9
    conv_tf = 2.0 * B * C * C * k * L / (ms_conv * 1e-3) / 1e12
    del x, out
    return {
        "shape": "[16,32,480000] (SpeechSR48 layer, batch 16; inputs > L2)",
        "act1d_kernel": {"bound": "hbm", "achieved": act_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": act_gbs / hbm_peak,
                         "ms": ms_act, "bytes": "4 B read + 2 B written per element"},
        "conv_umma_kernel": {"bound": "hbm", "achieved": conv_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": conv_gbs / hbm_peak, "ms": ms_conv, "tflops": conv_tf,
                             "bytes": "2 B operand + 4 B residual read + 4 B written per element (C=32, k=7: "
                                      "HBM-bound, SURVEY.md §7.3)"},
    }


def gpu_eager_baseline(wl, dev):
    """The reference's op sequence (oracle port: the same ATen calls as the reference modules) run EAGERLY by torch
    on the same B200 in strict fp32 (TF32 off) -- the second baseline BASELINE.md §3 promises.  Device-resident
    inputs, CUDA events, 2 warm-up + best of 5.  The oracle is used here as a timed BASELINE, never by the product."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        fwd = oracle_forward(wl, dev)
        with torch.no_grad():
            for _ in range(2):
                fwd()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fwd(); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
        return {"value": wl["audio_seconds"] / (best * 1e-3), "unit": "audio-s/s", "ms_per_step": best,
                "kind": "oracle port (reference ATen op sequence), torch eager on the same GPU, fp32, TF32 off, "
                        "cuDNN/ATen kernels", "sample": f"{wl['name']} (the full step), 2 warm-up + best of 5"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def run_config5(args, dev, world, rank):
    """BASELINE.json configs[4]: 512 utterances x 30 s (z [b,192,1500] -> wav [b,1,480000]), sharded by utterance
    over the ranks (runtime.shard_utterances), micro-batches of 32 equal-length utterances through one captured
    CUDA graph, peak-normalised int16 PCM on the device (to_pcm16), ONE device-side gather of the PCM onto rank 0
    (runtime.gather_waveforms -> torch.distributed.gather over NCCL).  Strong scaling: the job is fixed, time = max
    over ranks of (first H2D .. gather complete)."""
    import torch.distributed as dist

    import megatts2_hierspeechpp_b200 as hsv
    from megatts2_hierspeechpp_b200 import runtime as R
    from megatts2_hierspeechpp_b200 import synthetic as synth

    n_utt, T, mb = args.c5_utts, int(round(args.c5_seconds * 50)), args.c5_batch
    lengths = [T] * n_utt
    mine = R.shard_utterances(lengths, world, rank)
    batches = R.bucket_by_length(mine, lengths, mb)
    model = hsv.Vocoder()
    model.load_state_dict(synth.vocoder_sd(1234), strict=True)
    model.to(dev).eval()
    for m in model.modules():
        if hasattr(m, "parallel_blocks"):
            m.parallel_blocks = False          # the GPU is full at B=32 x 30 s: one stream
    runner = hsv.CudaGraphRunner(model)
    L = T * 320
    # synthetic inputs, generated per micro-batch on the device (seed = first utterance index) and parked in pinned
    # host memory: the timed loop starts from HOST buffers
    z_host, g_host = [], []
    for b in batches:
        gen = torch.Generator(device=dev).manual_seed(5000 + b[0])
        z_host.append(torch.randn(len(b), 192, T, generator=gen, device=dev).cpu().pin_memory())
        g_host.append(torch.randn(len(b), 256, 1, generator=gen, device=dev).cpu().pin_memory())
    z_d = torch.empty(mb, 192, T, device=dev)
    g_d = torch.empty(mb, 256, 1, device=dev)
    plan = [[i for b in R.bucket_by_length(R.shard_utterances(lengths, world, r), lengths, mb) for i in b]
            for r in range(world)]                       # every rank can derive every rank's pack order
    cap = max(len(p_) for p_ in plan)
    pcm = torch.zeros(cap, L, dtype=torch.int16, device=dev)   # one buffer per rank, padded to the largest shard
    with torch.no_grad():
        seen = set()
        for b, zh, gh in zip(batches, z_host, g_host):                  # capture every batch shape before timing
            if len(b) not in seen:
                seen.add(len(b))
                zb, gb = z_d[:len(b)], g_d[:len(b)]
                zb.copy_(zh); gb.copy_(gh)
                runner(zb, gb)
        R.to_pcm16(torch.zeros(2, 1, 64, device=dev), per_utterance=True)
        # NCCL sets up its point-to-point channels lazily at the first gather: do that outside the timed region
        if world > 1:
            R.gather_packed(torch.zeros(L, dtype=torch.int16, device=dev), dst=0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record()
        off = 0
        for b, zh, gh in zip(batches, z_host, g_host):
            n = len(b)
            zb, gb = z_d[:n], g_d[:n]
            zb.copy_(zh, non_blocking=True)
            gb.copy_(gh, non_blocking=True)
            wav = runner(zb, gb)
            pcm[off:off + n].copy_(R.to_pcm16(wav, per_utterance=True).view(n, L))
            off += n
        e1.record()
        recv = R.gather_packed(pcm.view(-1), dst=0) if world > 1 else pcm
        if recv is not None:
            recv = recv.view(world, cap, L)
        e2.record()
        local = {i: pcm[j] for j, i in enumerate(plan[rank])}
        merged = None if recv is None else {i: recv[r, j] for r in range(world) for j, i in enumerate(plan[r])}
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        compute_ms, total_ms = e0.elapsed_time(e1), e0.elapsed_time(e2)
        t = torch.tensor([compute_ms, total_ms, wall * 1e3], dtype=torch.float64, device=dev)
        tmin = t.clone()
        tsum = torch.tensor([len(mine) * T / 50.0 / (compute_ms * 1e-3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        # per-rank spot parity: this rank's first utterance, eager B=1 forward vs the reference op sequence in strict
        # fp32 on the same device (the oracle is the checker here, outside every timed region)
        from oracle import closed_form as CF
        from oracle import functional as OF
        i0 = batches[0][0]
        z1, g1 = z_host[0][:1].to(dev), g_host[0][:1].to(dev)
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            sd_dev = {k: v.to(dev) for k, v in synth.vocoder_sd(1234).items()}
            ref = OF.vocoder(sd_dev, z1, g1)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        got = model(z1, g1)
        ma, snr = CF.max_abs(ref.cpu().numpy(), got.cpu().numpy()), CF.snr_db(ref.cpu().numpy(), got.cpu().numpy())
        # and the PCM that went through the batched graph + gather equals the PCM of that eager forward
        pcm_ok = bool(torch.equal(R.to_pcm16(got, per_utterance=True).view(-1), local[i0]))
        par = torch.tensor([ma, -snr, 0.0 if pcm_ok else 1.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(par, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        complete = merged is not None and sorted(merged) == list(range(n_utt)) and \
            all(v.numel() == L and v.dtype == torch.int16 for v in merged.values())
        d2h_ms = None
        if merged:
            host = torch.empty(n_utt, L, dtype=torch.int16).pin_memory()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for i in range(n_utt):
                host[i].copy_(merged[i], non_blocking=True)
            d1.record()
            torch.cuda.synchronize()
            d2h_ms = d0.elapsed_time(d1)
        audio = n_utt * T / 50.0
        total_ms, comp_max = float(t[1]), float(t[0])
        out = {
            "workload": f"config #5: {n_utt} utterances x {args.c5_seconds:g} s, utterance-sharded over {world} GPU(s), "
                        f"micro-batch {mb}, int16 PCM gathered on rank 0 (device-side, NCCL)",
            "scaling": "strong", "n_gpus": world, "value": audio / (total_ms * 1e-3), "unit": "audio-s/s",
            "total_ms": total_ms, "wall_ms": float(t[2]), "compute_ms_max": comp_max, "compute_ms_min": float(tmin[0]),
            "gather_ms": total_ms - comp_max, "gather_frac": (total_ms - comp_max) / total_ms,
            "gather_bytes": int(n_utt * L * 2 * (world - 1) / max(1, world)),
            "sum_of_rank_rates": float(tsum[0]),
            "parallel_efficiency": (audio / (total_ms * 1e-3)) / float(tsum[0]),
            "parallel_efficiency_note": "job rate / sum over ranks of (rank audio / rank compute time): the cost of load "
                                        "imbalance + the gather; the driver computes scaling efficiency across N itself",
            "h2d_bytes": int(sum(z.numel() * 4 + g.numel() * 4 for z, g in zip(z_host, g_host)) * world),
            "d2h_ms_pcm_rank0": d2h_ms, "gathered_complete": bool(complete),
            "parity_spot": {"max_abs_worst_rank": float(par[0]), "snr_db_worst_rank": -float(par[1]),
                            "pcm_bit_exact_all_ranks": float(par[2]) == 0.0,
                            "what": "each rank's first utterance: eager forward vs the reference op sequence (torch fp32, "
                                    "TF32 off, same GPU); the gathered int16 PCM of that utterance equals the PCM of the "
                                    "eager forward bit for bit"},
        }
    del runner, model, pcm
    hsv.ops.clear_workspace()
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the B200 arm needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import megatts2_hierspeechpp_b200 as hsv
    from megatts2_hierspeechpp_b200 import _lib, build

    build.build()
    hbm_peak, tensor_peak, peak_kind = peaks()
    wl = make_workload(args, rank)
    model = build_model(wl, dev)
    if args.parallel_blocks:
        for m in model.modules():
            if hasattr(m, "parallel_blocks"):
                m.parallel_blocks = True
    host_in = [t.contiguous().pin_memory() for t in wl["host_inputs"]]
    dev_in = [t.to(dev) for t in host_in]
    runner = hsv.CudaGraphRunner(model) if not args.no_graph else None
    fwd = (lambda *a: runner(*a)) if runner else (lambda *a: model(*a))

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def flush():
        flush_buf.fill_(1.0)

    with torch.no_grad():
        n0 = _lib.LAUNCHES[0]
        model(*dev_in)                      # eager once: folds weights
        n0 = _lib.LAUNCHES[0]
        model(*dev_in)
        launches_per_step = _lib.LAUNCHES[0] - n0      # steady state (no fold/pack launches)
        out = fwd(*dev_in)
        static_in = runner.static_inputs(*dev_in) if runner else dev_in
        host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        torch.cuda.synchronize()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        h2d = sum(t.numel() * t.element_size() for t in host_in)
        d2h = host_out.numel() * host_out.element_size()

        def dev_step():
            fwd(*static_in)

        def e2e_step():
            for s_, h in zip(static_in, host_in):
                s_.copy_(h, non_blocking=True)
            o = fwd(*static_in)
            host_out.copy_(o, non_blocking=True)

        # the same end-to-end step with the step AFTER the path on the device: peak-normalise + int16
        # inside the captured graph, 2 bytes per sample over PCIe
        from megatts2_hierspeechpp_b200.runtime import to_pcm16
        pcm_runner = hsv.CudaGraphRunner(lambda *a: to_pcm16(model(*a), per_utterance=True)) if not args.no_graph else None
        pcm_out = pcm_runner(*dev_in) if pcm_runner else to_pcm16(model(*dev_in), per_utterance=True)
        pcm_static_in = pcm_runner.static_inputs(*dev_in) if pcm_runner else dev_in
        host_pcm = torch.empty(pcm_out.shape, dtype=torch.int16).pin_memory()

        def e2e_pcm_step():
            for s_, h in zip(pcm_static_in, host_in):
                s_.copy_(h, non_blocking=True)
            o = pcm_runner(*pcm_static_in) if pcm_runner else to_pcm16(model(*pcm_static_in), per_utterance=True)
            host_pcm.copy_(o, non_blocking=True)

        def timed_block(step):
            """EXACTLY args.steps steps, CUDA events around each (L2 flushed before each, outside the events)."""
            evs = []
            for _ in range(args.steps):
                flush()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); step(); e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            return [a.elapsed_time(b) for a, b in evs]

        # The timed block (exactly K steps) is repeated until the timed blocks span >= ~1 s of wall time, so that the
        # nvidia-smi sampler (100 ms period) sees the GPU under THIS load several times; the reported ms_per_step is
        # the median block (every block is K steps; `timed_blocks` says how many were run).
        for _ in range(args.warmup):
            flush(); dev_step()
        for _ in range(max(3, args.warmup)):
            e2e_step()
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        t_start = time.perf_counter()
        dev_blocks, e2e_blocks, pcm_blocks = [], [], []
        for _ in range(3):
            e2e_pcm_step()
        while True:
            barrier()
            dev_blocks.append(timed_block(dev_step))
            barrier()
            e2e_blocks.append(timed_block(e2e_step))
            barrier()
            pcm_blocks.append(timed_block(e2e_pcm_step))
            barrier()
            enough = time.perf_counter() - t_start >= args.min_seconds or len(dev_blocks) >= args.max_blocks
            flag = torch.tensor([1.0 if enough else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if float(flag) > 0:
                break
        clocks = sampler.stop()
        nb = len(dev_blocks)
        sums = torch.tensor([[sum(b) for b in dev_blocks], [sum(b) for b in e2e_blocks], [sum(b) for b in pcm_blocks]],
                            dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.MAX)     # max over ranks, block by block
        dev_ms = float(sums[0].median())
        e2e_ms = float(sums[1].median())
        pcm_ms = float(sums[2].median())
        step_ms = sorted(x for b in dev_blocks for x in b)

        roof = extra = shares = step_roof = cpu = sat = eager = None
        if rank == 0:
            par = [m for m in model.modules() if getattr(m, "parallel_blocks", False)]
            fams = algorithmic_work(model, dev_in)
            replays = cupti_step_profile(dev_step, flush) if not args.no_cupti else None
            roof, extra, shares, step_roof = kernel_roofline(fams, replays, dev_ms / args.steps, hbm_peak, tensor_peak,
                                                              peak_kind, wl["name"])
            try:
                sat = saturated_rooflines(dev, hbm_peak, tensor_peak)
            except Exception as e:  # e.g. not enough free memory next to a large workload
                sat = {"error": str(e)[:200]}
            if not args.no_gpu_eager:
                try:
                    eager = gpu_eager_baseline(wl, dev)
                except Exception as e:
                    eager = {"error": str(e)[:200]}
            if world == 1 and not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                cwl = wl
                sample = f"{wl['name']} (the full step)"
                if wl["audio_seconds"] > 40:      # bound the CPU sample to ~10-30 s of work
                    a2 = argparse.Namespace(**vars(args))
                    a2.batch, a2.seconds = 1, min(args.seconds, 10.0)
                    cwl = make_workload(a2, 0)
                    sample = f"{cwl['name']} (slice of the workload; CPU cost is linear in batch x duration)"
                fwd_cpu, kind = reference_forward(cwl)
                ts = time_cpu(fwd_cpu, 3, 1)
                cpu = {"value": cwl["audio_seconds"] / min(ts), "unit": "audio-s/s", "cores": cores, "kind": kind,
                       "sample": sample + (", the reference's own modules (baseline/_ref)" if kind == "reference" else
                                           ", oracle port of the reference's CPU path") + ", 1 warm-up + best of 3"}
        del runner, pcm_runner
        c5 = None
        if not args.no_config5:
            try:
                c5 = run_config5(args, dev, world, rank)
            except Exception as e:
                c5 = {"error": f"{type(e).__name__}: {e}"[:300]}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    total_audio = wl["audio_seconds"] * world * args.steps
    line = {
        "metric": "vocoder audio-sec/sec (RTF^-1)", "value": total_audio / (dev_ms / 1e3), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 residual stream + f16 tensor-core operands, f32 accumulate", "data": wl["data"],
        "config": {"workload": wl["name"], "per_gpu_batch": args.batch, "seconds_per_utterance": args.seconds,
                   "cuda_graph": not args.no_graph, "parallel_resblocks": bool(args.parallel_blocks),
                   "l2": "flushed (256 MB fill) before every timed step", "parallelism": f"utterance-sharded x{world}",
                   "timed_blocks": nb, "timing": f"{nb} blocks of exactly {args.steps} steps each (device-resident and "
                                                 "end-to-end blocks alternate); ms_per_step = median block / steps, max "
                                                 "over ranks per block",
                   "not_computed": "SourceNetwork.conv_post (the predicted f0 e_, 64->1 k7, < 0.01 % of the FLOPs): the "
                                   "vocoder path of SynthesizerTrn.infer discards it; Vocoder.forward skips it "
                                   "(need_pred=False) while the reference arm's sn(z, g) computes it"},
        "e2e": {"value": total_audio / (e2e_ms / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "e2e_pcm16": {"value": total_audio / (pcm_ms / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": int(host_pcm.numel() * 2), "ms_per_step": pcm_ms / args.steps,
                      "what": "e2e with the step after the path on the device: peak-normalised int16 PCM "
                              "(inference_plm.py:183-188) inside the graph, int16 D2H"},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "clocks": clocks, "roofline": roof, "roofline_other": extra, "step_roofline": step_roof,
        "roofline_saturated": sat, "kernel_shares": shares,
        "cpu_baseline": cpu, "gpu_eager_baseline": eager, "config5": c5,
        "step_ms_min_med_max": [step_ms[0], step_ms[len(step_ms) // 2], step_ms[-1]],
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="vocoder", choices=["vocoder", "speechsr48", "speechsr24", "chain24", "synth", "tts"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--parallel-blocks", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-cupti", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=1.2, help="repeat the K-step timed block until this much wall time")
    ap.add_argument("--max-blocks", type=int, default=400)
    ap.add_argument("--c5-utts", type=int, default=512)
    ap.add_argument("--c5-seconds", type=float, default=30.0)
    ap.add_argument("--c5-batch", type=int, default=32)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 5
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 50
        args.warmup = max(3, args.warmup if args.warmup is not None else 5)
        run_b200(args)


if __name__ == "__main__":
    main()
