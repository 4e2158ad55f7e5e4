"""Summarise a .ncu-rep (one profiled launch) as text for profiles/: key raw metrics + stall hot spots.

    python tools/ncu_summary.py gpurun_out/prof_act_sat.ncu-rep > profiles/r01_ncu_act_sat.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    print(f"# ncu summary of {path}")
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"kernel: {name}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {vals[i]:>18s} {units[i]}")
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    h = src[1]
    rows = src[2:]
    i_src, i_st, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    tot_s = sum(int(r[i_st] or 0) for r in rows) or 1
    tot_e = sum(int(r[i_ex] or 0) for r in rows) or 1
    print(f"\nSASS instructions: {len(rows)}, stall samples: {tot_s}, warp instructions executed: {tot_e}")
    print("top stall sites (share of samples, share of executed, SASS):")
    for i in sorted(range(len(rows)), key=lambda k: -int(rows[k][i_st] or 0))[:12]:
        r = rows[i]
        print(f"  #{i:5d} {100 * int(r[i_st] or 0) / tot_s:5.1f}% {100 * int(r[i_ex] or 0) / tot_e:5.1f}%  {r[i_src][:90]}")
    ops = {}
    for r in rows:
        t = r[i_src].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[i_ex] or 0)
    print("executed warp-instructions by opcode (top 14):")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:14]:
        print(f"  {op:10s} {100 * n / tot_e:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
