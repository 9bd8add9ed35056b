#!/usr/bin/env python3
"""Small mixed batch for compute-sanitizer (memcheck / racecheck) runs on the GPU box:
micro groups, cfg2 region groups, tandem-repeat groups that take the overflow pass, 400 bp
reads (128-anchor chain kernel), through lgr_genotype_batch and lgr_submit/lgr_wait, checked
against the oracle.  usage:
  compute-sanitizer --tool memcheck python tools/sanitize_case.py
  compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_case.py small"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle_lib as O  # noqa: E402
from compare import compare_results  # noqa: E402
from lancet2_b200 import abi, synth  # noqa: E402
from lancet2_b200.realign import GpuRealigner  # noqa: E402
from test_hostemu_parity import str_group  # noqa: E402


def main():
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    groups = synth.make_groups(21, 2 if small else 4, n_reads=48, n_haps=4, hap_len=700)
    groups += synth.make_region_groups(9, ref_len=20_000)[:2 if small else 5]
    groups += [str_group(np.random.default_rng(70), 150, 800, 3, 24 if small else 64)]
    groups += synth.make_groups(5, 1, read_len=400, hap_len=1500, n_haps=3, n_reads=16 if small else 48, sub_err=0.01)
    batch = abi.Batch(groups)
    gpu = GpuRealigner(0)
    want, _ = O.oracle_genotype(batch, gpu.params, n_threads=8)
    got, st = gpu.genotype_batch(batch)
    errs = compare_results(batch, want, got)
    t1, r1 = gpu.submit(batch)
    t2, r2 = gpu.submit(batch)
    gpu.wait(t2), gpu.wait(t1)
    errs += compare_results(batch, want, r1) + compare_results(batch, want, r2)
    # the packed wire format (unpack kernels) and the long-read point (banded extension, > 64 anchors per pair)
    pk, _ = gpu.genotype_packed(abi.PackedBatch(groups, gpu.lib), batch)
    errs += compare_results(batch, want, pk)
    long_groups = synth.make_groups(1000 * 250 + 1000 + 8, 1, read_len=250, hap_len=1000, n_haps=4, n_reads=24 if small else 96)
    lb = abi.Batch(long_groups)
    lw, _ = O.oracle_genotype(lb, gpu.params, n_threads=8)
    lg, _ = gpu.genotype_packed(abi.PackedBatch(long_groups, gpu.lib), lb)
    errs += compare_results(lb, lw, lg)
    gpu.close()
    # the two other entry-point families: repeat scan, AddToTable + FORMAT math on the device
    import repeat_lib
    from lancet2_b200.format_metrics import GpuFormatMetrics
    from lancet2_b200.repeat_scan import GpuRepeatScan
    jobs = repeat_lib.window_jobs(3, 4 if small else 12, k_values=(13, 31, 64), lengths=(300, 700))
    rs = GpuRepeatScan(0)
    rep, _ = rs.scan(jobs)
    orc = repeat_lib.oracle()
    errs += [f"repeat job {i}" for i, (sq, k, mm) in enumerate(jobs) if int(rep[i]) != orc.orc_has_repeat(sq, len(sq), k, mm)]
    rs.close()
    rng = np.random.default_rng(2)
    nr = batch.n_reads
    fmt = GpuFormatMetrics(0)
    from test_gpu_evidence_build import variant_tables
    kk, vlen = variant_tables(groups, batch)
    recs, keys, _ = fmt.from_assign(batch, n_samples=3, sample_id=rng.integers(0, 3, nr).astype(np.int32),
                                    start0=rng.integers(0, 10_000, nr).astype(np.int64), isize=rng.integers(-500, 500, nr).astype(np.int64),
                                    sam_flag=rng.integers(0, 64, nr).astype(np.uint16), mapq=rng.integers(0, 61, nr).astype(np.uint8),
                                    softclip=rng.integers(0, 2, nr).astype(np.uint8), var_n_alleles=kk, var_len=vlen,
                                    host_assign=got.assign[:batch.n_assign])
    fmt.close()
    print(f"sanitize_case: repeat jobs {len(jobs)}, supports {len(recs)}")
    print(f"sanitize_case: {batch.n_pairs} pairs, {st.kernel_launches} launches, mismatches: {len(errs)}")
    sys.exit(1 if errs else 0)


if __name__ == "__main__":
    main()
