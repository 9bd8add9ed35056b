#!/usr/bin/env python3
"""Turn an ncu report of bench.py (cfg2) into the tracked summaries under profiles/:
  profiles/<tag>_ncu_cfg2.txt     raw metrics per kernel + per-source-line attribution
  profiles/<tag>_ncu_static.json  per-launch DRAM bytes / warp-instructions of k_chain_warp (bench.py reads r2's)
usage: tools/make_profile_summary.py gpurun_out/<report>.ncu-rep [tag] [int_peak.json]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_config_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
    peak_file = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r1_int_peak.json")
    raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, units = raw[0], raw[1]
    out, txt = {}, ["ncu --set full --import-source on --clock-control none -c 13, one resident pass over the cfg2 workload (233,282 pairs; tools/debug_counters.py cfg2, no L2 flush before the pass)", ""]
    for r in raw[2:]:
        short = r[h.index("Kernel Name")].split("(")[0].split("::")[-1]
        d = {k: r[h.index(k)] for k in KEYS if k in h}
        out.setdefault(short, {k: float(v) for k, v in d.items()})
        txt.append(f"== {short} ==")
        txt += [f"  {k} = {d[k]} {units[h.index(k)] if h.index(k) < len(units) else ''}" for k in KEYS if k in d]
        txt.append("")
    lib = os.path.join(ROOT, "lancet2_b200", "csrc", "liblancet_gpu_realign.so")
    for mangled, human in (("k_chain_warpILi64", "k_chain_warp"), ("k_chain_coldILi64", "k_chain_cold"), ("k_ext_warp", "k_ext_warp"),
                           ("k_finish_warp", "k_finish_warp"), ("k_read_sketch", "k_read_sketch"), ("k_assign", "k_assign")):
        txt.append(f"==== per-source-line attribution: {human} ====")
        txt.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, lib, mangled, "25", human],
                                  capture_output=True, text=True).stdout)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_cfg2.txt"), "w") as fh:
        fh.write("\n".join(txt) + "\n")
    # bench.py's ms_k_map spans the hot and the cold chain kernel: both go into the per-launch constants
    chains = [v for k, v in out.items() if k.startswith("k_chain_warp") or k.startswith("k_chain_cold")]
    peak = json.load(open(peak_file))
    static = {"cfg2": {"kernel": "k_chain_warp + k_chain_cold",
                       "dram_bytes_per_launch": sum(ch["dram__bytes_read.sum"] + ch["dram__bytes_write.sum"] for ch in chains) * 1e6,
                       "warp_inst_per_launch": sum(ch["smsp__inst_executed.sum"] for ch in chains),
                       "source": f"profiles/{tag}_ncu_cfg2.txt (ncu --set full, one launch)"},
              "int_issue_peak_warp_inst_per_s": peak["iadd3"] * 1e12 / 32,
              "int_issue_peak_source": "profiles/r1_int_peak.json (tools/int_peak.cu, IADD3 thread-ops/s / 32)"}
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_static.json"), "w") as fh:
        json.dump(static, fh, indent=1)
    for k, v in out.items():
        print(k, v["gpu__time_duration.sum"], "ms", int(v["smsp__inst_executed.sum"]), "inst", v["smsp__issue_active.avg.pct_of_peak_sustained_active"], "% issue")


if __name__ == "__main__":
    main()
