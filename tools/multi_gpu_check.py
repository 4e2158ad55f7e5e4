"""Utterance-sharded inference on N GPUs (torchrun): every rank runs its shard of a seeded utterance set
through the vocoder, rank 0 gathers all waveforms (NCCL barrier + gather = the only collectives) and checks
them bit-identical against the same utterances run alone on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megatts2_hierspeechpp_b200 as hsv  # noqa: E402
from megatts2_hierspeechpp_b200.runtime import bucket_by_length, gather_waveforms, shard_utterances  # noqa: E402
from megatts2_hierspeechpp_b200 import synthetic as synth  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = hsv.Vocoder()
    m.load_state_dict(synth.vocoder_sd(1234), strict=True)
    m.to(dev).eval()
    n_utt = 12
    lengths = [20 + 10 * (i % 3) for i in range(n_utt)]          # frames (0.4 - 0.8 s)
    inputs = [synth.vocoder_inputs(1, lengths[i], seed=100 + i) for i in range(n_utt)]
    mine = shard_utterances(lengths, world, rank)
    out = {}
    with torch.no_grad():
        for mb in bucket_by_length(mine, lengths, max_batch=4):
            z = torch.cat([inputs[i][0] for i in mb]).to(dev)
            g = torch.cat([inputs[i][1] for i in mb]).to(dev)
            wav = m(z, g)
            for j, i in enumerate(mb):
                out[i] = wav[j:j + 1].clone()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    allw = gather_waveforms(out, dst=0)
    if rank == 0:
        assert sorted(allw) == list(range(n_utt)), sorted(allw)
        with torch.no_grad():
            for i in range(n_utt):
                ref = m(inputs[i][0].to(dev), inputs[i][1].to(dev)).cpu()
                assert ref.shape == (1, 1, lengths[i] * 320)
                assert torch.equal(ref, allw[i].cpu()), f"utterance {i} differs between sharded and single runs"
        print(f"MULTI_GPU_OK world={world} utterances={n_utt}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
