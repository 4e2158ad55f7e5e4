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
torch.cuda.synchronize()
assert torch.equal(w, w2) and torch.equal(w, w3), "variants disagree"
print("sanitize target ok", tuple(w.shape), int(pcm.abs().max()))
