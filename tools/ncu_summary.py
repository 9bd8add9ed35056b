#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + SASS source page) into a short text report.
usage: tools/ncu_summary.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, vals = raw[0], raw[2] if len(raw) > 2 else raw[1]
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
            "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "sass__inst_executed_local_loads"]
    units = raw[1] if len(raw) > 2 else [""] * len(hdr)
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h} = {v} {u}")
    sass = list(csv.reader(run(["ncu", "-i", rep, "--page", "source", "--csv"]).splitlines()))
    h2 = sass[1]
    ci = {h: i for i, h in enumerate(h2)}
    data = []
    for n, r in enumerate(sass[2:]):
        try:
            data.append((n, float(r[ci["# Samples"]]), float(r[ci["Instructions Executed"]]),
                         float(r[ci["Avg. Threads Executed"]] or 0), r[ci["Source"]].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[1] for d in data) or 1
    toti = sum(d[2] for d in data) or 1
    hist = collections.Counter()
    for d in data:
        hist[min(int(d[3]) // 8 * 8, 32)] += d[2]
    print("instructions by active-lane bucket (%):", sorted((k, round(100 * v / toti, 1)) for k, v in hist.items()))
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    # contiguous hot regions: cumulative samples per 40-instruction window
    win = 40
    regions = []
    for st in range(0, len(data), win):
        chunk = data[st:st + win]
        regions.append((sum(c[1] for c in chunk), st, sum(c[2] for c in chunk), chunk))
    regions.sort(reverse=True)
    print(f"top regions ({win}-instruction windows):")
    for samp, st, inst, chunk in regions[:top_n]:
        hot = max(chunk, key=lambda c: c[1])
        thr = sum(c[3] * c[2] for c in chunk) / max(1, sum(c[2] for c in chunk))
        print(f"  sass[{st:5d}..] {100 * samp / tot:5.1f}% samples, {100 * inst / toti:5.1f}% inst, avg lanes {thr:4.1f} | hottest: {hot[4][:70]}")


if __name__ == "__main__":
    main()
