#!/usr/bin/env python3
"""The fixed corpus for pinning the minimap2 restatement against the real minimap2 2.30
(tools/capture_mm2_golden.c, tests/test_mm2_golden.py): SURVEY.md §8(d)'s seeds 42 / 137 / 271, the
shapes the parity suite uses (exact reads, SNV/indel haplotypes, overhangs, tandem repeats with
high-occurrence seeds and > 64 anchors, reverse-complemented and N-containing reads, 250 / 600 bp reads).

  python tools/export_mm2_cases.py cases  > cases.txt     input of capture_mm2_golden
  python tools/export_mm2_cases.py oracle > oracle.jsonl  what THIS repository's oracle returns, in the
                                                          capture program's output format (diff-able)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from lancet2_b200 import abi, synth  # noqa: E402


def corpus():
    """deterministic list of groups; ASCII only, no spaces in names"""
    from test_hostemu_parity import str_group
    groups = []
    for seed in (42, 137, 271):
        groups += synth.make_groups(seed, 4, read_len=150, hap_len=900, n_haps=4, n_reads=64)
        groups += synth.make_groups(seed + 1, 2, read_len=250, hap_len=1500, n_haps=4, n_reads=32, sub_err=0.01, indel_err=0.002)
        groups += synth.make_groups(seed + 2, 2, read_len=100, hap_len=300, n_haps=2, n_reads=48)
        groups += synth.make_region_groups(seed, ref_len=30_000)[:4]
        groups += [str_group(np.random.default_rng(seed + 3), 150, 800, 3, 48)]
        groups += [str_group(np.random.default_rng(seed + 4), 250, 1000, 3, 24, (1, 4), (30, 120))]
        groups += [str_group(np.random.default_rng(seed + 5), 600, 1500, 2, 12, (2, 30), (5, 40))]
    return groups


def write_cases(out):
    for g in corpus():
        out.write(f"G {len(g.haps)} {len(g.reads)}\n")
        for h in g.haps:
            out.write("H " + h.decode() + "\n")
        for nm, r in zip(g.names, g.reads):
            out.write(f"R {nm} {r.decode()}\n")


def oracle_records(groups, mid_occ_per_group=None):
    """the oracle's answer in the capture program's record format (dicts)"""
    import oracle_lib as O
    prm = O.default_params()
    if mid_occ_per_group is not None:
        for g, m in zip(groups, mid_occ_per_group):
            g.mid_occ = int(m)
    batch = abi.Batch(groups)
    res, _ = O.oracle_genotype(batch, prm, n_threads=os.cpu_count() or 1)
    return records_from_result(batch, groups, res)


def records_from_result(batch, groups, res):
    recs = []
    for gi, g in enumerate(groups):
        for r in range(len(g.reads)):
            rg = int(batch.grp_read_begin[gi]) + r
            for h in range(len(g.haps)):
                pair = int(batch.pair_off[rg]) + h
                a = res.aln[pair]
                d = {"g": gi, "r": r, "h": h, "n_regs": int(a["n_regs"])}
                if a["valid"]:
                    d.update(score=int(a["score"]), rs=int(a["rs"]), re=int(a["re"]), qs=int(a["qs"]), qe=int(a["qe"]), rev=int(a["rev"]),
                             mlen=int(a["mlen"]), blen=int(a["blen"]), dp_score=int(a["dp_score"]), dp_max=int(a["dp_max"]),
                             n_ambi=int(a["n_ambi"]), cigar=res.cigar(pair))
                recs.append(d)
    return recs


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "cases"
    if mode == "cases":
        write_cases(sys.stdout)
    elif mode == "oracle":
        for d in oracle_records(corpus()):
            sys.stdout.write(json.dumps(d, separators=(",", ":")) + "\n")
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
