#!/usr/bin/env python3
"""First measurement of the FORMAT-math row: lgr_format_metrics on the supports a cfg2
step produces (253 variants x 2 samples, ≈ 230 evidence records each, mates deduplicated) and on a
deep-coverage shape (≈ 2000 records per support).  Prints one JSON line per shape: CUDA-event time of
k_fmt_dedup + k_fmt_metrics, the wall time of the whole call from host buffers, and — on the
box's host cores, one thread — the same arithmetic compiled by g++ (tests/hostemu, checker only) and the
reference's own VariantSupport (oracle/_ref, when it was built)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import format_lib as F  # noqa: E402
from lancet2_b200 import abi  # noqa: E402
from lancet2_b200.format_metrics import GpuFormatMetrics  # noqa: E402


def reference_all_cores(sups, cores):
    """the reference's VariantSupport over the supports, split over `cores` threads (ctypes drops the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    chunks = [abi.EvidenceBatch(sups[i::cores]) for i in range(cores) if sups[i::cores]]
    F.ref_format_batch(chunks[0])  # warm
    with ThreadPoolExecutor(len(chunks)) as ex:
        t0 = time.perf_counter()
        list(ex.map(F.ref_format_batch, chunks))
        return (time.perf_counter() - t0) * 1e3


def main():
    fmt = GpuFormatMetrics(0)
    rng = np.random.default_rng(42)
    quick = os.environ.get("LGR_FMT_QUICK") == "1"  # profiler runs: the cfg2-step shape only, few repetitions
    shapes = (("cfg2-step", 506, 230), ("deep", 506, 2000))
    for name, n_sup, n_rec in shapes[:1] if quick else shapes:
        sups = [F.random_support(rng, n=int(rng.integers(int(n_rec * 0.8), int(n_rec * 1.2))), n_alleles=2, dup_frac=0.4)
                for _ in range(n_sup)]
        batch = abi.EvidenceBatch(sups)
        for _ in range(3):
            got, _ = fmt.compute(batch)
        ms_k, wall = [], []
        for _ in range(2 if quick else 20):
            t0 = time.perf_counter()
            got, ms = fmt.compute(batch)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms_k.append(ms)
        t0 = time.perf_counter()
        rc, want = F.emu_format(batch)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        errs = F.compare_format(want, got)
        ref_ms = ref_all_ms = None
        cores = os.cpu_count() or 1
        if F.have_ref():  # the reference's own VariantSupport (oracle/_ref): one host thread, then all of them
            t0 = time.perf_counter()
            ref_rec = F.ref_format_batch(batch)
            ref_ms = (time.perf_counter() - t0) * 1e3
            errs += F.compare_format(ref_rec, got, label="vs reference: ")
            ref_all_ms = reference_all_cores(sups, cores)
        bytes_in = sum(v.nbytes for v in batch.cols.values())
        # same roofline object as bench.py: algorithmic bytes of the call over the kernels' time against the measured
        # HBM peak — it documents that the row is NOT HBM bound; `issue` is the bound that matters (profiles/r1_format_ncu.txt)
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            peak, peak_src = 6650.0, "fallback 6650 GB/s"
        k_ms = float(np.median(ms_k))
        achieved = (bytes_in + got.nbytes) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": "k_fmt_dedup + k_fmt_metrics", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": 5.16e6 + 0.38e6 if name == "cfg2-step" else None, "peak_source": peak_src,
                    "note": "issue bound: k_fmt_metrics runs at 0.58 of the measured INT issue peak on the cfg2-step shape "
                            "(ncu, profiles/r1_format_ncu.txt); traffic = dram bytes of that capture"}
        print(json.dumps({"workload": name, "supports": n_sup, "evidence_records": batch.n_evidence,
                          "ms_kernels_median": float(np.median(ms_k)), "ms_call_median": float(np.median(wall)),
                          "supports_per_s_kernels": n_sup / (float(np.median(ms_k)) * 1e-3),
                          "h2d_bytes": int(bytes_in), "d2h_bytes": int(got.nbytes),
                          "cpu_same_arithmetic_ms_1thread": cpu_ms, "cpu_reference_ms_1thread": ref_ms,
                          "cpu_reference_ms_all_cores": ref_all_ms, "cores": cores,
                          "mismatches_vs_host_build": len(errs), "roofline": roofline}))


if __name__ == "__main__":
    main()
