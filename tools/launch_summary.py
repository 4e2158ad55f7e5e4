"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.

    python tools/launch_summary.py gpurun_out/launches_forward.csv [last_n_forwards]
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
    return rows


def short(name):
    for key in ("act1d_kernel", "conv_umma_persist_kernel", "conv_umma_kernel", "conv1d_tiled_kernel", "conv1d_thin_kernel",
                "conv_transpose1d_kernel", "sr_pre_interp_kernel", "pack_weight_kernel", "weight_norm_fold_kernel",
                "pack_blk16_kernel", "nearest_gather_kernel", "add3_bcast_kernel", "interp_table_kernel"):
        if key in name:
            if key == "act1d_kernel":
                return key + ("<blk16>" if "ELi1E" in name or ", 1>" in name else "<f32>")
            return key
    return name[:60]


if __name__ == "__main__":
    rows = load(sys.argv[1])
    if "--all" in sys.argv:
        # whole command (e.g. bench.py: warm-up + timed steps + per-kernel roofline passes): every hsv launch
        # except the one-time weight fold / pack kernels and torch's own fill/copy kernels
        sel = [r for r in rows if "fold" not in r[0] and "pack_weight" not in r[0] and "at::" not in r[0]
               and "Memcpy" not in r[0]]
    else:
        # every forward issues the same launches apart from the one-time weight fold / pack kernels: drop those
        # and keep the last of the nfwd forwards
        steady = [r for r in rows if "fold" not in r[0] and "pack_weight" not in r[0]]
        nfwd = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1
        sel = steady[-(len(steady) // nfwd):]
    agg = defaultdict(lambda: [0, 0.0])
    for name, us, *_ in sel:
        a = agg[short(name)]
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"launches={len(sel)} total={tot:.1f} us")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:34s} n={n:4d} total={us:9.1f} us  avg={us / n:8.2f} us  share={us / tot:6.3f}")
    if "-v" in sys.argv:
        for name, us, g, b in sel:
            print(f"{short(name):34s} {us:9.2f} us grid={g} block={b}")
