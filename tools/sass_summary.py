"""Per-kernel SASS evidence of the shipped library (profiles/sass_summary.txt): counts of the Blackwell tensor / TMA
mnemonics (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, cp.async -> LDGSTS, fma.rn.f32x2 ->
FFMA2) and of the legacy tensor path (HMMA must be 0).  Runs on CPU: cuobjdump -sass on libhsv.so.

    python tools/sass_summary.py > profiles/sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "megatts2_hierspeechpp_b200", "libhsv.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "MUFU", "HMMA",
        "HGMMA", "BAR", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)[:70]
            cur = per.setdefault(name, dict.fromkeys(KEYS, 0))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for k in KEYS[:-1]:
                if op.startswith(k):
                    cur[k] += 1
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass megatts2_hierspeechpp_b200/libhsv.so (sm_100a), HEAD {head}: SASS mnemonic counts per kernel")
    print(f"{'kernel':72s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = dict.fromkeys(KEYS, 0)
    for name, c in per.items():
        print(f"{name:72s} " + " ".join(f"{c[k]:7d}" for k in KEYS))
        for k in KEYS:
            tot[k] += c[k]
    print(f"{'ALL KERNELS':72s} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor instructions found"


if __name__ == "__main__":
    main()
