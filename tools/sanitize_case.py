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
    gpu.close()
    print(f"sanitize_case: {batch.n_pairs} pairs, {st.kernel_launches} launches, mismatches: {len(errs)}")
    sys.exit(1 if errs else 0)


if __name__ == "__main__":
    main()
