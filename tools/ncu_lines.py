#!/usr/bin/env python3
"""Attribute an ncu source-page (SASS) profile to CUDA source lines using nvdisasm -g line info.
usage: tools/ncu_lines.py report.ncu-rep lib.so mangled_kernel_substring [top_n] [human_kernel_substring]
The library must be the same build the report was taken from."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, kern = sys.argv[1:4]
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    start, dis = None, []
    for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):  # one per translation unit
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
        for i, l in enumerate(dis):
            if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"):
                start = i
                break
        if start is not None:
            break
    assert start is not None, "kernel not found"
    line_of = []  # per instruction index → (file, line)
    cur = ("?", 0)
    for l in dis[start + 1:]:
        if l.startswith("//---------------------") or (l.startswith(".text.") and l.rstrip().endswith(":")):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            line_of.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    parts = out.split('"Kernel Name",')[1:]
    human = sys.argv[5] if len(sys.argv) > 5 else None
    part = parts[0]
    if human:
        part = [p for p in parts if human in p.split("\n", 1)[0]][0]
    sass = list(csv.reader(('"Kernel Name",' + part).splitlines()))
    h2 = sass[1]
    ci = {h: i for i, h in enumerate(h2)}
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    local = collections.defaultdict(lambda: [0.0, 0.0])  # executed STL / LDL per source line (LGR_NCU_LOCAL=1)
    src_col = ci.get("Source")
    n = 0
    tot_s = tot_i = 0.0
    for r in sass[2:]:
        try:
            samp, inst, thr = float(r[ci["# Samples"]]), float(r[ci["Instructions Executed"]]), float(r[ci["Thread Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        key = line_of[n] if n < len(line_of) else ("?", -1)
        a = agg[key]
        a[0] += samp; a[1] += inst; a[2] += thr
        if src_col is not None and src_col < len(r):
            if " STL" in " " + r[src_col]:
                local[key][0] += inst
            elif " LDL" in " " + r[src_col]:
                local[key][1] += inst
        tot_s += samp; tot_i += inst
        n += 1
    print(f"instructions in profile {n}, in disassembly {len(line_of)}")
    if os.environ.get("LGR_NCU_LOCAL") == "1":
        print(f"local memory: {sum(v[0] for v in local.values()):.0f} STL, {sum(v[1] for v in local.values()):.0f} LDL executed")
        for (f, ln), (st, ld) in sorted(local.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:top_n]:
            print(f"  STL {st:10.0f}  LDL {ld:10.0f}  {f}:{ln}")
        return
    src_cache = {}
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]
    print(" %inst  %samp  lanes  file:line  source")
    for (f, ln), (s, i, t) in rows:
        if f not in src_cache:
            p = os.path.join(os.path.dirname(os.path.abspath(lib)), f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src_cache[f][ln - 1].strip()[:80] if 0 < ln <= len(src_cache[f]) else ""
        print(f"{100*i/tot_i:6.2f} {100*s/tot_s:6.2f} {t/max(i,1):6.1f}  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
