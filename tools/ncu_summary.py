#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key throughput metrics, stall breakdown, hottest source lines.
usage: tools/ncu_summary.py <file.ncu-rep> [--lines N]"""
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    return {n: (v[i], units[i]) for i, n in enumerate(h)}


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    m = raw(rep)
    print("kernel:", m.get("Kernel Name", ("?",))[0], " grid", m.get("Grid Size", ("?",))[0], " block", m.get("Block Size", ("?",))[0])
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__inst_executed.sum", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg",
            "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
    for k in keys:
        if k in m:
            print(f"  {k:72s} {m[k][0]:>18s} {m[k][1]}")
    stalls = []
    for k, (v, u) in m.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            x = fnum(v)
            if x:
                stalls.append((x, k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    print("  stall reasons (warps stalled per issue-active cycle):")
    for x, k in sorted(stalls, reverse=True)[:10]:
        print(f"    {k:40s} {x:8.2f}")
    # source page: hottest lines by sampled stalls
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if not rows:
        return
    h = rows[0]
    try:
        i_src = h.index("Source")
        i_smp = next(i for i, n in enumerate(h) if n.startswith("# Samples") or n == "Warp Stall Sampling (All Samples)" or n.startswith("Warp Stall Sampling (All"))
    except (ValueError, StopIteration):
        print("  (no source page)", h[:12])
        return
    i_inst = next((i for i, n in enumerate(h) if n.startswith("Instructions Executed")), None)
    i_line = next((i for i, n in enumerate(h) if n in ("#", "Line", "Address")), 0)
    tot = sum(fnum(r[i_smp]) or 0 for r in rows[1:] if len(r) > i_smp)
    top = sorted((r for r in rows[1:] if len(r) > i_smp), key=lambda r: -(fnum(r[i_smp]) or 0))[:nlines]
    print(f"  hottest lines by stall samples (total {tot:.0f}):")
    for r in top:
        s = fnum(r[i_smp]) or 0
        ie = r[i_inst] if i_inst is not None else ""
        print(f"    {100 * s / max(tot, 1):5.1f}%  inst={ie:>12s}  {r[i_line]:>6s}  {r[i_src].strip()[:110]}")


if __name__ == "__main__":
    main()
