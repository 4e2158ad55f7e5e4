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
    # the fp32 attention kernel, the multi-stream mode (resblock streams, two WaveNet stacks side by side, consumer-side
    # resblock sums), the tail of the text-to-vec model
    hsv.ops.set_mha_variant(1)
    o1 = syn.voice_conversion_noise_control(w2v.to(dev), torch.LongTensor([6]).to(dev), mel.to(dev),
                                            torch.LongTensor([10, 7]).to(dev), f0.to(dev))
    hsv.ops.set_mha_variant(0)
    par = [mm for mm in list(syn.modules()) + list(m.modules()) if hasattr(mm, "parallel_blocks")]
    for mm in par:
        mm.parallel_blocks = True
    torch.manual_seed(3)
    o2 = syn.voice_conversion_noise_control(w2v.to(dev), torch.LongTensor([6]).to(dev), mel.to(dev),
                                            torch.LongTensor([10, 7]).to(dev), f0.to(dev))
    w5 = m(z.to(dev), g.to(dev))
    for mm in par:
        mm.parallel_blocks = False
    torch.manual_seed(3)
    o3 = syn.voice_conversion_noise_control(w2v.to(dev), torch.LongTensor([6]).to(dev), mel.to(dev),
                                            torch.LongTensor([10, 7]).to(dev), f0.to(dev))
    tail = hsv.TTVTail()
    tail.load_state_dict(synth.ttv_tail_sd(3456), strict=True)
    tail.to(dev).eval()
    zt, mt, gt = synth.ttv_tail_inputs(2, 9, seed=4, lengths=[9, 5])
    w2v_t, pitch_t = tail(zt.to(dev), mt.to(dev), gt.to(dev))
    sines, uv = hsv.ops.sinegen(torch.rand(2, 9, device=dev) * 300, 320, 16000.0, 4)
    buf = hsv.ops.blk16_buffer(1, 16, 64, dev, slot=9)
    hsv.ops.pack_blk16(torch.randn(1, 16, 64, device=dev), buf)
    st = hsv.ops.blk16_stats(buf, 16, 64)
torch.cuda.synchronize()
assert (w4 - w).abs().max().item() < 2e-3 and bool(torch.isfinite(o).all()) and int(st[0]) == 0
assert torch.equal(w, w2) and torch.equal(w, w3) and torch.equal(w, w5) and torch.equal(o2, o3), "variants disagree"
assert bool(torch.isfinite(o1).all()) and bool(torch.isfinite(pitch_t).all()) and tuple(pitch_t.shape) == (2, 1, 36)
print("sanitize target ok", tuple(w.shape), int(pcm.abs().max()))
