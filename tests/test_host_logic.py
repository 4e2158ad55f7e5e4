"""Host-side logic: state_dict layout, config, sharding (incl. a world_size-2 gloo run)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
import megatts2_hierspeechpp_b200 as hsv
from megatts2_hierspeechpp_b200.runtime import bucket_by_length, shard_utterances
from oracle import synth


def _manifest(name):
    return [s.split(":") for s in golden("state_dict_manifest.npz")[name].tolist()]


@pytest.mark.parametrize("name,factory", [
    ("dec", lambda: hsv.Generator(**hsv.HIER_CFG)),
    ("sn", lambda: hsv.SourceNetwork(256)),
    ("sr", lambda: hsv.SpeechSR24(100, 40, **hsv.SR_CFG)),
])
def test_state_dict_matches_reference_manifest(name, factory):
    m = factory()
    sd = m.state_dict()
    man = _manifest(name)
    assert [k for k, _ in man] == list(sd.keys())
    for k, shp in man:
        assert ",".join(map(str, sd[k].shape)) == shp, k


def test_synthetic_checkpoints_load_strict():
    v = hsv.Vocoder()
    v.load_state_dict(synth.vocoder_sd(1234), strict=True)
    hsv.SpeechSR48(128, 40, **hsv.SR_CFG).load_state_dict(synth.speechsr_sd(), strict=True)
    n_dec = sum(p.numel() for p in v.dec.parameters())
    n_sn = sum(p.numel() for p in v.sn.parameters())
    assert n_dec == 15165152 and n_sn == 2432384      # SURVEY.md §0.4 [measured on the reference]


def test_real_checkpoint_loads_and_filters_check():
    sd = {k: torch.from_numpy(v.copy()) for k, v in golden("speechsr24_state.npz").items()}
    m = hsv.SpeechSR24(100, 40, **hsv.SR_CFG)
    m.load_state_dict(sd, strict=True)
    m.dec.activation_post.check_filters()
    with torch.no_grad():
        m.dec.activation_post.upsample.filter[0, 0, 3] += 1e-3
    with pytest.raises(ValueError, match="constant-folds"):
        m.dec.activation_post.check_filters()


def test_kaiser_sinc_filter_matches_checkpoint_taps():
    f = hsv.kaiser_sinc_filter1d(0.25, 0.3, 12).flatten()
    assert (f - torch.tensor(hsv.modules.FILTER_TAPS)).abs().max() < 5e-8
    assert hsv.get_padding(11, 5) == 25 and hsv.get_padding(3, 1) == 1


def test_cpu_forward_fails_loudly():
    m = hsv.SpeechSR24(100, 40, **hsv.SR_CFG)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 64))


def test_sharding_is_a_balanced_partition():
    rng = np.random.RandomState(0)
    lengths = rng.randint(100, 1500, size=101).tolist()
    for ws in (1, 2, 4, 8):
        parts = [shard_utterances(lengths, ws, r) for r in range(ws)]
        assert sorted(i for p in parts for i in p) == list(range(101))
        counts = [len(p) for p in parts]
        assert max(counts) - min(counts) <= 1
        tot = [sum(lengths[i] for i in p) for p in parts]
        assert max(tot) - min(tot) <= 1500
    b = bucket_by_length(list(range(10)), [5, 5, 5, 7, 7, 5, 5, 5, 5, 7], 4)
    assert all(len(set([5, 5, 5, 7, 7, 5, 5, 5, 5, 7][i] for i in mb)) == 1 and len(mb) <= 4 for mb in b)
    assert sorted(i for mb in b for i in mb) == list(range(10))


def test_two_rank_gloo_shard_and_gather(tmp_path):
    """world_size-2 job on CPU/gloo: each rank takes its shard, 'processes' it, rank 0 gathers everything."""
    script = tmp_path / "job.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from megatts2_hierspeechpp_b200.runtime import shard_utterances, gather_waveforms\n"
        "dist.init_process_group('gloo')\n"
        "r, ws = dist.get_rank(), dist.get_world_size()\n"
        "lengths = [10 + (7 * i) % 13 for i in range(9)]\n"
        "mine = shard_utterances(lengths, ws, r)\n"
        "local = {i: torch.full((1, lengths[i] * 320), float(i)) for i in mine}\n"
        "dist.barrier()\n"
        "allw = gather_waveforms(local, dst=0)\n"
        "if r == 0:\n"
        "    assert sorted(allw) == list(range(9)), sorted(allw)\n"
        "    assert all(allw[i].shape == (1, lengths[i] * 320) and float(allw[i][0, 0]) == i for i in allw)\n"
        "    print('GATHER_OK', len(allw))\n"
        "else:\n"
        "    assert allw is None\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK 9" in r.stdout


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm) prints ONE JSON line with the
    contract's keys; run here on a 0.2 s utterance so it takes seconds."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--seconds", "0.2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["unit"] == "audio-s/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0


def test_two_rank_gloo_gather_packed(tmp_path):
    """The metadata-free collective under gather_waveforms: equal-size int16 buffers, raw-byte transport."""
    script = tmp_path / "job2.py"
    script.write_text(
        "import sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from megatts2_hierspeechpp_b200.runtime import gather_packed\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "flat = (torch.arange(1000, dtype=torch.int32) * (r + 1) % 30000).to(torch.int16)\n"
        "out = gather_packed(flat, dst=0)\n"
        "if r == 0:\n"
        "    assert out.shape == (2, 1000) and out.dtype == torch.int16\n"
        "    for q in range(2):\n"
        "        assert torch.equal(out[q], (torch.arange(1000, dtype=torch.int32) * (q + 1) % 30000).to(torch.int16))\n"
        "    print('PACKED_OK')\n"
        "else:\n"
        "    assert out is None\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29619", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PACKED_OK" in r.stdout


def test_gate_permutation_groups_tanh_and_sigmoid_halves():
    """The channel order the operand-writing gate epilogue expects of a WN in_layer (2H outputs): consecutive groups of
    [8 tanh | 8 sigmoid] channels, i.e. unit j of 16 columns holds channels 8j..8j+7 and H+8j..H+8j+7 -- a permutation."""
    from megatts2_hierspeechpp_b200 import ops
    for H in (192, 512, 8):
        perm = ops.gate_permutation(2 * H)
        assert sorted(perm.tolist()) == list(range(2 * H))
        g = perm.view(H // 8, 2, 8)
        assert torch.equal(g[:, 0], torch.arange(H).view(H // 8, 8))
        assert torch.equal(g[:, 1], H + torch.arange(H).view(H // 8, 8))
    with pytest.raises(ValueError):
        ops.gate_permutation(2 * 12)


def test_folded_gate_and_cat_weights_follow_their_layers():
    """_FoldedGate permutes weight and bias together (and per layer block for cond_layer); _FoldedCat concatenates the
    adaLN layers of a flow block in module order and refreshes when a parameter changes."""
    import megatts2_hierspeechpp_b200.front as FR
    from megatts2_hierspeechpp_b200 import ops
    torch.manual_seed(0)
    H, n = 16, 3
    conv = torch.nn.Conv1d(8, 2 * H * n, 1)
    fg = FR._FoldedGate(conv, blocks=n)
    w, b = fg.weight(), fg.bias()
    base = ops.gate_permutation(2 * H)
    perm = torch.cat([base + i * 2 * H for i in range(n)])
    assert torch.equal(w, conv.weight.detach()[perm]) and torch.equal(b, conv.bias.detach()[perm])
    lins = [torch.nn.Linear(4, 6) for _ in range(3)]
    fc = FR._FoldedCat(lins)
    assert torch.equal(fc.weight()[:, :, 0], torch.cat([l.weight for l in lins]).detach())
    assert torch.equal(fc.bias(), torch.cat([l.bias for l in lins]).detach())
    with torch.no_grad():
        lins[1].bias.add_(1.0)
    assert torch.equal(fc.bias(), torch.cat([l.bias for l in lins]).detach())
