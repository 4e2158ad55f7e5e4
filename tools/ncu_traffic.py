"""DRAM traffic per launch of the two hot kernel families, from an ncu metrics CSV of one workload:

    ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
        --log-file gpurun_out/traffic_vocoder.csv python tools/profile_kernels.py forward
    python tools/ncu_traffic.py <workload name> gpurun_out/traffic_vocoder.csv [n_forwards]   # -> profiles/ncu_traffic.json

bench.py reads profiles/ncu_traffic.json to fill `roofline.traffic` (bytes per launch, averaged like `achieved`).
ncu flushes the caches before every profiled kernel, so these are cold-cache figures: reads = compulsory
operand traffic, writes only what leaves L2 during the launch."""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    per_id = defaultdict(dict)
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3,
                 "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
        per_id[int(r["ID"])]["name"] = r["Kernel Name"]
        per_id[int(r["ID"])][r["Metric Name"]] = v * scale
    return [per_id[k] for k in sorted(per_id)]


def main(workload, path, nfwd=1):
    rows = load(path)
    # every forward issues the same launches apart from the one-time weight fold / pack kernels: drop those and
    # keep the last forward
    steady = [r for r in rows if "fold" not in r["name"] and "pack_weight" not in r["name"]]
    per = len(steady) // nfwd
    sel = steady[-per:]
    fam = {"act1d": [r for r in sel if "act1d_kernel" in r["name"]],
           "conv1d_umma": [r for r in sel if "conv_umma" in r["name"]]}
    out = {}
    for k, rs in fam.items():
        if not rs:
            continue
        rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in rs)
        wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in rs)
        out[k] = {"launches": len(rs), "dram_bytes_per_launch": (rd + wr) / len(rs), "dram_read_per_launch": rd / len(rs),
                  "dram_write_per_launch": wr / len(rs),
                  "avg_launch_us_under_ncu": sum(r.get("gpu__time_duration.sum", 0.0) for r in rs) / len(rs) / 1e3,
                  "source": os.path.basename(path)}
    dst = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    allw = {}
    if os.path.isfile(dst):
        with open(dst) as f:
            allw = json.load(f)
    allw[workload] = out
    import datetime
    allw["_meta"] = {"head": os.environ.get("HSV_HEAD", "unknown"), "when": datetime.datetime.utcnow().isoformat() + "Z",
                     "how": "ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum (cold cache per "
                            "kernel) over one eager forward; tools/ncu_traffic.py"}
    with open(dst, "w") as f:
        json.dump(allw, f, indent=1, sort_keys=True)
    print(json.dumps({workload: out}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
