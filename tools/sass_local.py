#!/usr/bin/env python3
"""Where a kernel touches local memory: STL/LDL instructions of its SASS by source line (nvdisasm -g).
usage: tools/sass_local.py lib.so kernel_name_substring"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
    tmp = tempfile.mkdtemp(dir=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"))
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
        inside, cur, n_inst = False, None, 0
        st, ld = collections.Counter(), collections.Counter()
        for l in sass:
            if l.startswith(".text."):
                inside = kern in l
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            if re.search(r"^\s+/\*[0-9a-f]{4}\*/", l):
                n_inst += 1
            if re.search(r"\bSTL(\.\w+)*\b", l):
                st[cur] += 1
            if re.search(r"\bLDL(\.\w+)*\b", l):
                ld[cur] += 1
        if n_inst:
            print(f"{cubin}: {kern}: {n_inst} instructions, {sum(st.values())} STL, {sum(ld.values())} LDL")
            for name, c in (("STL", st), ("LDL", ld)):
                for k, v in c.most_common(12):
                    print(f"  {name} {k[0]}:{k[1]}  x{v}")


if __name__ == "__main__":
    main()
