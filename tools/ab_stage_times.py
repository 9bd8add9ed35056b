#!/usr/bin/env python3
"""A/B helper: median per-stage kernel times (CUDA events inside the library) of resident passes over
one or more workloads, for the library LGR_LIBRARY points at.
usage: [LGR_LIBRARY=lancet2_b200/csrc/variants/libX.so] python tools/ab_stage_times.py cfg2 l250"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lancet2_b200 import abi  # noqa: E402
from lancet2_b200.realign import GpuRealigner  # noqa: E402


def main():
    names = sys.argv[1:] or ["cfg2"]
    gpu = GpuRealigner(0)
    for name in names:
        groups, _ = bench._build_replica_workload(name, 42)
        packed = abi.PackedBatch(groups, gpu.lib)
        gpu.upload_packed(packed)
        rows = []
        for it in range(25):
            st = gpu.run_resident()
            if it >= 5:
                rows.append((st.ms_k_index, st.ms_k_sketch, st.ms_k_map, st.ms_k_ext, st.ms_k_assign, st.ms_kernels))
        med = np.median(np.array(rows), axis=0)
        print(json.dumps({"lib": os.environ.get("LGR_LIBRARY", "default"), "workload": name, "pairs": int(st.n_pairs),
                          "ms": dict(zip(["index", "sketch", "map", "ext_finish", "assign", "all"], [round(float(x), 4) for x in med])),
                          "M_pairs_per_s": round(st.n_pairs / med[5] / 1e3, 2),
                          "arena_MiB": round(gpu.lib.lgr_arena_bytes(gpu._ctx) / 2**20, 1)}), flush=True)
    gpu.close()


if __name__ == "__main__":
    main()
