"""Property tests of the minimap2 restatement (oracle/mm2_restate.cpp; PARITY
UNPINNED — no reference test or binary pins mm_map).  These are
the self-made known-answer cases listed in SURVEY.md §8(c)(i)."""
import numpy as np
import pytest

import oracle_lib as O
from lancet2_b200 import abi, synth


def rand_hap(seed, n=1000):
    return synth._rand_bases(np.random.default_rng(seed), n).tobytes()


def test_exact_substring_is_LM():
    hap = rand_hap(1)
    for st in (0, 1, 7, 300, 850):
        d = O.map_debug(hap, hap[st:st + 150])
        a = d["aln"]
        assert a["valid"] == 1 and d["cigar_str"] == "150M"
        assert (a["rs"], a["re"], a["qs"], a["qe"]) == (st, st + 150, 0, 150)
        assert a["dp_max"] == 150 and a["mlen"] == 150 and a["blen"] == 150


def test_single_snv_is_LM_with_one_mismatch():
    hap = rand_hap(2)
    r = bytearray(hap[200:350])
    r[75] = ord("A") if r[75] != ord("A") else ord("C")
    d = O.map_debug(hap, bytes(r))
    assert d["cigar_str"] == "150M" and d["aln"]["mlen"] == 149 and d["aln"]["dp_max"] == 145


def test_deletion_is_left_shifted():
    # hap has a homopolymer run; deleting inside it must left-align the D
    rng = np.random.default_rng(3)
    left, right = synth._rand_bases(rng, 400).tobytes(), synth._rand_bases(rng, 400).tobytes()
    left = left[:-1] + b"C"
    right = b"G" + right[1:]
    hap = left + b"AAAAAAAA" + right
    read = left[-70:] + b"AAAAA" + right[:75]  # 3-base deletion inside the A-run
    d = O.map_debug(hap, read)
    assert d["cigar_str"] == "70M3D80M", d["cigar_str"]
    assert d["aln"]["rs"] == 330


def test_insertion_and_overhang_softclip():
    hap = rand_hap(4)
    r = hap[300:370] + b"ACGTTGCA" + hap[370:442]
    d = O.map_debug(hap, r)
    ops = d["cigar_str"]
    assert "8I" in ops and d["aln"]["qs"] == 0 and d["aln"]["qe"] == 150
    # read hanging 10 bases over the left end of the haplotype → leading I is stripped into qs
    d = O.map_debug(hap, b"GATTACAGAT" + hap[0:140])
    assert d["aln"]["qs"] == 10 and d["aln"]["rs"] == 0 and d["cigar_str"] == "140M"


def test_unrelated_read_has_no_hit():
    d = O.map_debug(rand_hap(5), rand_hap(6)[:150])
    assert d["n_regs"] == 0 and d["aln"]["valid"] == 0


def test_reverse_complement_read_maps_on_reverse_strand():
    hap = rand_hap(7)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    r = hap[100:250].translate(comp)[::-1]
    d = O.map_debug(hap, r)
    assert d["aln"]["valid"] == 1 and d["aln"]["rev"] == 1 and d["cigar_str"] == "150M"
    assert (d["aln"]["rs"], d["aln"]["re"]) == (100, 250)


def test_radix_sort_matches_stable_sort_below_65_and_is_a_permutation_above():
    # chain on a long read (> 64 anchors) still yields a consistent alignment
    hap = rand_hap(8, 1500)
    d = O.map_debug(hap, hap[100:1100])
    assert len(d["anchors"]) > 64
    xs = [x for x, _ in d["anchors"]]
    assert xs == sorted(xs)
    assert d["cigar_str"] == "1000M"


def test_batch_oracle_runs_and_assigns_reads():
    groups = synth.make_groups(11, 3, n_reads=64, n_haps=4, hap_len=600)
    batch = abi.Batch(groups)
    res, st = O.oracle_genotype(batch)
    assert st.n_pairs == batch.n_pairs and st.n_aligned > 0.9 * batch.n_pairs
    asg = res.assign[:batch.n_assign]
    assert asg["assigned"].sum() > 0
    # an assigned record points at a haplotype that carries the allele it reports
    for g_i, g in enumerate(groups):
        r0, v0 = batch.grp_read_begin[g_i], batch.grp_var_begin[g_i]
        V = len(g.variants)
        for r in range(len(g.reads)):
            for v in range(V):
                a = asg[batch.asg_off[r0 + r] + v]
                if a["assigned"]:
                    s, l, al = g.variants[v][a["hap_id"]]
                    assert al == a["allele"]
    # threads do not change results
    res2, _ = O.oracle_genotype(batch, n_threads=4)
    assert (res.aln[:batch.n_pairs] == res2.aln[:batch.n_pairs]).all()
    assert res.assign[:batch.n_assign].tobytes() == res2.assign[:batch.n_assign].tobytes()
