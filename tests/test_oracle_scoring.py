"""Pins the Lancet-owned half of the oracle (oracle/genotype_oracle.cpp):
 (1) the 11 known-answer cases of the reference's tests/hts/cigar_utils_test.cpp:58-172,
 (2) golden vectors produced by the reference's own sources compiled unmodified
     (tests/golden/scoring_golden.json, made by tests/golden/gen_scoring_golden.py),
 (3) the Phred LUT and ENCODE_TABLE, all 256 entries, bit for bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O

OPC = {c: i for i, c in enumerate("MIDNSHP=XB")}
GOLD = os.path.join(os.path.dirname(__file__), "golden", "scoring_golden.json")


def cig(s):
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | OPC[ch])
            num = ""
    return np.asarray(out, dtype=np.uint32)


def enc(s):
    lib = O.load_oracle()
    return np.asarray([lib.orc_lancet_encode(c) for c in s.encode()], dtype=np.uint8)


def nm(c, q, t):
    lib = O.load_oracle()
    cg, qq, tt = cig(c), enc(q), enc(t)
    return lib.orc_edit_distance(cg.ctypes.data, cg.size, qq.ctypes.data, qq.size, tt.ctypes.data, tt.size)


def r2q(c, pos):
    lib = O.load_oracle()
    cg = cig(c)
    return lib.orc_refpos_to_qpos(cg.ctypes.data, cg.size, pos)


# reference: tests/hts/cigar_utils_test.cpp:58-127
@pytest.mark.parametrize("c,q,t,want", [
    ("10M", "ACGTACGTAC", "ACGTACGTAC", 0),
    ("5M", "ATGTA", "AAGAA", 2),
    ("3M2I3M", "ACGTTACG", "ACGACG", 2),
    ("3M2D3M", "ACGACG", "ACGTTACG", 2),
    ("3S5M2S", "NNNACGTANN", "ACGTA", 0),
    ("4M1I2M1D4M", "ACATCACTAAA", "ACGTACGTAAA", 3),
    ("3=1X2=", "ACGTAC", "ACGAAC", 1),
])
def test_edit_distance_known_answers(c, q, t, want):
    assert nm(c, q, t) == want


# reference: tests/hts/cigar_utils_test.cpp:132-172
@pytest.mark.parametrize("c,pos,want", [
    ("10M", 0, 0), ("10M", 5, 5), ("10M", 9, 9),
    ("3M2I5M", 2, 2), ("3M2I5M", 3, 5), ("3M2I5M", 7, 9),
    ("3M2D5M", 2, 2), ("3M2D5M", 3, 3), ("3M2D5M", 4, 3), ("3M2D5M", 5, 3), ("3M2D5M", 6, 4),
    ("3S5M", 0, 3), ("3S5M", 4, 7),
])
def test_refpos_to_qpos_known_answers(c, pos, want):
    assert r2q(c, pos) == want


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as fh:
        return json.load(fh)


def test_phred_and_encode_tables(gold):
    lib = O.load_oracle()
    for q in range(256):
        assert float(lib.orc_phred_err(q)).hex() == gold["phred"][q], q
        assert lib.orc_lancet_encode(q) == gold["encode"][q], q
    assert lib.orc_phred_err(1000) == lib.orc_phred_err(255)


def test_scoring_against_reference_golden(gold):
    lib = O.load_oracle()
    for n, c in enumerate(gold["cases"]):
        cg = np.asarray(c["cigar"], dtype=np.uint32)
        hap = np.asarray(c["hap"], dtype=np.uint8)
        read = np.asarray(c["read"], dtype=np.uint8)
        quals = np.asarray(c["quals"], dtype=np.uint8)
        tgt = hap[c["rs"]:c["re"]].copy()
        e = c["expect"]
        o3 = np.zeros(3)
        bq = C.c_uint8(0)
        lib.orc_local_score(cg.ctypes.data, cg.size, read.ctypes.data, read.size, tgt.ctypes.data, tgt.size,
                            quals.ctypes.data, quals.size, c["rs"], c["var_start"], c["var_len"], o3.ctypes.data,
                            C.byref(bq))
        assert (o3[0].hex(), o3[1].hex(), o3[2].hex(), bq.value) == (e["pbq"], e["raw"], e["identity"], e["min_bq"]), n
        assert lib.orc_edit_distance(cg.ctypes.data, cg.size, read.ctypes.data, read.size, tgt.ctypes.data,
                                     tgt.size) == e["nm"], n
        assert lib.orc_softclip_penalty(cg.ctypes.data, cg.size) == e["sc_pen"], n
        got = [lib.orc_refpos_to_qpos(cg.ctypes.data, cg.size, rp) for rp in (0, 1, 5, 17, 60, 10000)]
        assert got == e["qpos"], n
        # combined: global_score = (i32)((score - sc_pen) - raw)   (combined_scorer.cpp:74-78)
        gs = int(np.trunc((float(c["score"]) - e["sc_pen"]) - float.fromhex(e["raw"])))
        assert gs == e["global_score"], n
        comb = float(gs) + float.fromhex(e["pbq"]) * float.fromhex(e["identity"])
        assert comb.hex() == e["combined"], n
