"""Host-side runtime around the kernels: CUDA-graph replay, utterance sharding, reference patching.

The path has no collective: every utterance is independent, so multi-GPU
inference is one process per GPU, each running its shard; ``torch.distributed`` is used only for
the final gather of the waveforms and the timing barrier.
"""
from __future__ import annotations

import sys
from typing import Callable, Dict, List, Sequence, Tuple

import torch


def to_pcm16(audio: torch.Tensor, scale_norm: str = "max", prompt_audio_max: float = 1.0, order: str = "plm",
             per_utterance: bool = False) -> torch.Tensor:
    """int16 PCM of a generated waveform, on the device (the step after the path).

    ``order="plm"``: ``audio / |audio|.max() * 32767.0 * s`` with s = 0.999, or ``prompt_audio_max`` when
    ``scale_norm == "prompt"`` (inference_plm.py:183-188); ``order="speechsr"``: ``... * 0.999 * 32767.0``
    (inference_speechsr.py:39-41).  Bit-exact with the reference's fp32 arithmetic + ``astype('int16')``."""
    from . import ops
    if order == "plm":
        s1, s2 = 32767.0, (float(prompt_audio_max) if scale_norm == "prompt" else 0.999)
    elif order == "speechsr":
        s1, s2 = 0.999, 32767.0
    else:
        raise ValueError("order must be 'plm' or 'speechsr'")
    pcm, _ = ops.peak_norm_pcm16(audio.detach().contiguous(), s1, s2, per_row=per_utterance)
    return pcm


class CudaGraphRunner:
    """Capture ``fn(*tensors)`` once per input-shape signature and replay it.

    One forward of the vocoder is 288 small kernels; replaying them as a CUDA graph removes the
    launch gaps that dominate batch-1 latency.  Inputs are copied into static buffers; the returned
    tensors are the graph's static outputs (valid until the next call with the same shapes)."""

    def __init__(self, fn: Callable, warmup: int = 2):
        self.fn = fn
        self.warmup = warmup
        self._graphs: Dict[Tuple, Tuple] = {}

    def _key(self, args):
        return tuple((tuple(a.shape), a.dtype, a.device.index) for a in args)

    def __call__(self, *args: torch.Tensor):
        for a in args:
            if not a.is_cuda:
                raise RuntimeError("CudaGraphRunner: CUDA tensors only")
        key = self._key(args)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = [torch.empty_like(a) for a in args]
            for s, a in zip(static_in, args):
                s.copy_(a)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(self.warmup):  # folds weights, allocates workspaces, loads modules
                    self.fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                static_out = self.fn(*static_in)
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        for s, a in zip(static_in, args):
            if s.data_ptr() != a.data_ptr():
                s.copy_(a, non_blocking=True)
        graph.replay()
        return static_out

    def static_inputs(self, *args):
        """The static input buffers for this shape signature (capture first if needed)."""
        self(*args)
        return self._graphs[self._key(args)][1]


def shard_utterances(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank ``rank`` processes: sort by length (longest first), deal
    round-robin in a serpentine order so every rank gets the same count (+-1) and near-equal total
    length.  Deterministic; the union over ranks is a partition of range(len(lengths))."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    mine = []
    for pos, idx in enumerate(order):
        rnd, slot = divmod(pos, world_size)
        owner = slot if rnd % 2 == 0 else world_size - 1 - slot
        if owner == rank:
            mine.append(idx)
    return mine


def bucket_by_length(indices: Sequence[int], lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Micro-batches of equal-length utterances (the Generator takes no masks, and replicate padding
    inside Activation1d makes padded batching differ near the boundary: SURVEY.md §8e)."""
    by_len: Dict[int, List[int]] = {}
    for i in indices:
        by_len.setdefault(int(lengths[i]), []).append(i)
    out = []
    for L in sorted(by_len, reverse=True):
        ids = by_len[L]
        for s in range(0, len(ids), max_batch):
            out.append(ids[s:s + max_batch])
    return out


def gather_waveforms(local: Dict[int, torch.Tensor], dst: int = 0):
    """Final gather of {utterance index: waveform} onto rank ``dst`` (the only collective of the job)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    payload = {k: v.cpu() for k, v in local.items()}
    gathered = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return merged


def patch_reference(modules=None) -> List[str]:
    """Swap the reference's hot-path classes for the B200 ones in already-imported reference modules
    (``hierspeechpp_speechsynthesizer``, ``speechsr24k.speechsr``, ``speechsr48k.speechsr``,
    ``alias_free_torch``), so the reference's own ``SynthesizerTrn``/inference scripts build and call
    them.  Returns the list of patched attributes.  See INTEGRATION.md."""
    from . import modules as M

    mods = modules if modules is not None else sys.modules
    patched = []

    def _set(modname, attr, obj):
        m = mods.get(modname)
        if m is not None:
            setattr(m, attr, obj)
            patched.append(f"{modname}.{attr}")

    for attr, obj in (("Generator", M.Generator), ("SourceNetwork", M.SourceNetwork), ("AMPBlock1", M.AMPBlock1),
                      ("DBlock", M.DBlock), ("Activation1d", M.Activation1d)):
        _set("hierspeechpp_speechsynthesizer", attr, obj)
    _set("speechsr24k.speechsr", "Generator", M.SpeechSR24Generator)
    _set("speechsr24k.speechsr", "AMPBlock0", M.AMPBlock0)
    _set("speechsr24k.speechsr", "Activation1d", M.Activation1d)
    _set("speechsr48k.speechsr", "Generator", M.SpeechSR48Generator)
    _set("speechsr48k.speechsr", "AMPBlock0", M.AMPBlock0)
    _set("speechsr48k.speechsr", "Activation1d", M.Activation1d)
    for attr, obj in (("Activation1d", M.Activation1d), ("UpSample1d", M.UpSample1d),
                      ("DownSample1d", M.DownSample1d), ("LowPassFilter1d", M.LowPassFilter1d)):
        _set("alias_free_torch", attr, obj)
    _set("activations", "SnakeBeta", M.SnakeBeta)
    return patched
