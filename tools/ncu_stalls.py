"""Per-instruction stall reasons of one profiled launch (.ncu-rep with --import-source on).

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    body = rows[2:]
    i_src, i_all, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    reasons = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[i_all] or 0) for r in body) or 1
    agg = {n: sum(int(r[i] or 0) for r in body) for i, n in reasons}
    print(f"# {path}: {len(body)} SASS instructions, {tot} stall samples")
    print("stall reasons (all instructions): " + ", ".join(f"{n[6:]}={100 * v / tot:.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    print("top sites: idx  %samples  exec  dominant reasons  SASS")
    for k in sorted(range(len(body)), key=lambda k: -int(body[k][i_all] or 0))[:top]:
        r = body[k]
        rs = sorted(((int(r[i] or 0), n[6:]) for i, n in reasons), reverse=True)[:2]
        print(f"  #{k:5d} {100 * int(r[i_all] or 0) / tot:5.1f}% {r[i_ex]:>8s}  " + " ".join(f"{n}:{v}" for v, n in rs if v) + f"   {r[i_src][:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
