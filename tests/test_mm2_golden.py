"""Pins the minimap2 restatement (oracle/mm2_restate.cpp) and the CUDA path against the REAL minimap2
2.30 — when a golden capture exists.  tests/golden/mm2_golden.jsonl.gz is produced by
tools/capture_mm2_golden.c on a machine that has minimap2 (this repository and its build container do
not: see the file's header); until someone commits it these tests SKIP and every parity claim of the
repository carries the caveat "minimap2 half unpinned" (oracle/MM2_AUDIT.md).
The corpus itself (tools/export_mm2_cases.py) is checked here unconditionally."""
import gzip
import io
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import export_mm2_cases as X  # noqa: E402

from lancet2_b200 import abi  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "mm2_golden.jsonl.gz")
# fields of mm_reg1_t / mm_extra_t the capture prints and the path reproduces ("cnt" is not carried by lgr_aln)
FIELDS = ["n_regs", "score", "rs", "re", "qs", "qe", "rev", "mlen", "blen", "dp_score", "dp_max", "n_ambi", "cigar"]


def load_golden():
    if not os.path.exists(GOLD):
        pytest.skip("no real-minimap2 capture committed (tools/capture_mm2_golden.c): minimap2 half stays unpinned")
    mids, recs = {}, []
    with gzip.open(GOLD, "rt") as fh:
        for line in fh:
            d = json.loads(line)
            if "mid_occ" in d:
                mids[d["g"]] = d["mid_occ"]
            else:
                recs.append(d)
    return mids, recs


def diff(want, got):
    bad = []
    assert len(want) == len(got)
    for w, g in zip(want, got):
        assert (w["g"], w["r"], w["h"]) == (g["g"], g["r"], g["h"])
        for f in FIELDS:
            if w.get(f) != g.get(f):
                bad.append(f"g{w['g']} r{w['r']} h{w['h']} {f}: minimap2 {w.get(f)} ours {g.get(f)}")
    return bad


def test_corpus_is_deterministic_and_well_formed():
    buf = io.StringIO()
    X.write_cases(buf)
    text = buf.getvalue()
    buf2 = io.StringIO()
    X.write_cases(buf2)
    assert text == buf2.getvalue()
    lines = text.splitlines()
    n_g = sum(1 for l in lines if l.startswith("G "))
    assert n_g == len(X.corpus()) >= 40
    i = 0
    while i < len(lines):
        _, nh, nr = lines[i].split()
        nh, nr = int(nh), int(nr)
        assert all(l.startswith("H ") and set(l[2:]) <= set("ACGTN") for l in lines[i + 1:i + 1 + nh])
        assert all(l.startswith("R ") and len(l.split(" ")) == 3 for l in lines[i + 1 + nh:i + 1 + nh + nr])
        i += 1 + nh + nr


def test_oracle_equals_real_minimap2():
    mids, want = load_golden()
    groups = X.corpus()
    got = X.oracle_records(groups, [mids[g] for g in range(len(groups))])
    bad = diff(want, got)
    assert not bad, "\n".join(bad[:40])


@pytest.mark.gpu
def test_cuda_path_equals_real_minimap2():
    mids, want = load_golden()
    from lancet2_b200.realign import GpuRealigner
    groups = X.corpus()
    for g, grp in enumerate(groups):
        grp.mid_occ = int(mids[g])
    batch = abi.Batch(groups)
    gpu = GpuRealigner(0)
    try:
        res, _ = gpu.genotype_batch(batch, arena=1 << 22)
    finally:
        gpu.close()
    bad = diff(want, X.records_from_result(batch, groups, res))
    assert not bad, "\n".join(bad[:40])
