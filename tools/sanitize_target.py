"""Small end-to-end target for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, tiny shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402
import megatts2_hierspeechpp_b200.modules as M  # noqa: E402
from megatts2_hierspeechpp_b200 import synthetic as synth  # noqa: E402

dev = "cuda:0"
m = hsv.Vocoder()
m.load_state_dict(synth.vocoder_sd(1234), strict=True)
m.to(dev).eval()
z, g = synth.vocoder_inputs(2, 6, seed=5)
with torch.no_grad():
    w = m(z.to(dev), g.to(dev))
    M.FUSE_MAX_CHANNELS[0] = 64          # activation-producing conv variant
    w2 = m(z.to(dev), g.to(dev))
    M.FUSE_MAX_CHANNELS[0] = 0
    hsv.ops.set_umma_debug(128)          # persistent variant wherever it can run
    w3 = m(z.to(dev), g.to(dev))
    hsv.ops.set_umma_debug(0)
    pcm = hsv.to_pcm16(w)
    # tensor-core activation variant, frame-rate front, sine source, operand statistics
    from megatts2_hierspeechpp_b200 import _lib
    _lib.load().hsv_set_act_variant(2)
    w4 = m(z.to(dev), g.to(dev))
    _lib.load().hsv_set_act_variant(0)
    syn = hsv.HierSpeechSynthesizer()
    syn.load_state_dict(synth.synthesizer_sd(1234), strict=True)
    syn.to(dev).eval()
    w2v, f0, mel = synth.synthesizer_inputs(6, 10, seed=2)
    o = syn.voice_conversion_noise_control(w2v.to(dev), torch.LongTensor([6]).to(dev), mel.to(dev),
                                           torch.LongTensor([10, 7]).to(dev), f0.to(dev))
    sines, uv = hsv.ops.sinegen(torch.rand(2, 9, device=dev) * 300, 320, 16000.0, 4)
    buf = hsv.ops.blk16_buffer(1, 16, 64, dev, slot=9)
    hsv.ops.pack_blk16(torch.randn(1, 16, 64, device=dev), buf)
    st = hsv.ops.blk16_stats(buf, 16, 64)
torch.cuda.synchronize()
assert (w4 - w).abs().max().item() < 2e-3 and bool(torch.isfinite(o).all()) and int(st[0]) == 0
assert torch.equal(w, w2) and torch.equal(w, w3), "variants disagree"
print("sanitize target ok", tuple(w.shape), int(pcm.abs().max()))
