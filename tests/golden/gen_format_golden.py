#!/usr/bin/env python3
"""Generate tests/golden/format_golden.json with the REFERENCE's own VariantSupport
(src/lancet/caller/variant_support.cpp, genotype_likelihood.cpp, posterior_base_qual.cpp and
base/mann_whitney.h compiled unmodified into oracle/_ref by `make -C oracle ref`):
  * `random`: seeded evidence streams and every FORMAT accessor's value (lgr_format records, hex);
  * `edge`: tests/format_lib.py:edge_supports (extreme qualities and magnitudes, ties, single alleles);
  * `scipy`: the rows of the reference's scipy-derived Mann-Whitney fixture
    (tests/data/base/mann_whitney_scipy_ref.tsv, consumed by tests/base/mann_whitney_test.cpp:231-330)
    — inputs and expected effect sizes, read from the reference tree at generation time.
Run in the build container only (needs /root/reference and oracle/_ref)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import format_lib as F  # noqa: E402
from lancet2_b200 import abi  # noqa: E402

TSV = "/root/reference/tests/data/base/mann_whitney_scipy_ref.tsv"


def to_json(sup):
    return {k: (np.asarray(v).tolist() if not np.isscalar(v) else int(v)) for k, v in sup.items()}


def main():
    rng = np.random.default_rng(2024)
    sups = [F.random_support(rng) for _ in range(110)]
    sups += [F.random_support(rng, n=int(rng.integers(150, 400)), n_alleles=2) for _ in range(6)]
    sups += [F.random_support(rng, n=60, n_alleles=8), F.random_support(rng, n=0, n_alleles=2)]
    ref = F.ref_format(sups)
    edge = F.edge_supports()
    ref_edge = F.ref_format(edge)
    scipy_rows = []
    with open(TSV) as fh:
        next(fh)
        for line in fh:
            c = line.rstrip("\n").split("\t")
            if len(c) < 6:
                continue
            exp = float(c[5])  # nan = an empty group: the reference returns nullopt
            scipy_rows.append({"ref": [float(x) for x in c[3].split(",") if x], "alt": [float(x) for x in c[4].split(",") if x],
                               "expected": None if exp != exp else exp})
    json.dump({"source": "reference VariantSupport compiled unmodified (oracle/_ref); scipy rows from the reference's "
                         "tests/data/base/mann_whitney_scipy_ref.tsv",
               "dtype_itemsize": abi.FORMAT_DTYPE.itemsize,
               "random": [{"support": to_json(s), "record": ref[i:i + 1].tobytes().hex()} for i, s in enumerate(sups)],
               "edge": [{"support": to_json(s), "record": ref_edge[i:i + 1].tobytes().hex()} for i, s in enumerate(edge)],
               "scipy": scipy_rows},
              open(os.path.join(HERE, "format_golden.json"), "w"), separators=(",", ":"))
    print("wrote", len(sups), "random supports and", len(scipy_rows), "scipy rows")


if __name__ == "__main__":
    main()
