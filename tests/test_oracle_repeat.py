"""SURVEY.md §8f #3 (first half) — pins oracle/repeat_oracle.cpp: the reference's known-answer tests
(reference tests/base/repeat_test.cpp) and, where oracle/_ref holds it, the reference's own
src/lancet/base/repeat.cpp compiled unmodified, on seeded windows."""
import numpy as np
import pytest

import repeat_lib

ORC = repeat_lib.oracle()
REF = repeat_lib.reference()
LIBS = [("orc", ORC)] + ([("ref", REF)] if REF is not None else [])


def _ham(lib, prefix, a, b):
    assert len(a) == len(b)
    return getattr(lib, prefix + "_hamming_dist")(a, b, len(a))


def _rep_kmers(lib, prefix, kmers, mm):
    k = len(kmers[0]) if kmers else 4
    return bool(getattr(lib, prefix + "_has_repeat_kmers")(b"".join(kmers), len(kmers), k, mm))


@pytest.mark.parametrize("prefix,lib", LIBS)
def test_hamming_known_answers(prefix, lib):
    # repeat_test.cpp:76-86 (small), :98-158 (SIMD width boundaries, single byte, empty)
    assert _ham(lib, prefix, b"aaaa", b"aaaa") == 0
    assert _ham(lib, prefix, b"aaaa", b"abaa") == 1
    assert _ham(lib, prefix, b"aaaa", b"aaba") == 1
    assert _ham(lib, prefix, b"abaa", b"aaba") == 2
    assert _ham(lib, prefix, b"A" * 32, b"A" * 32) == 0
    assert _ham(lib, prefix, b"A" * 32, b"C" * 32) == 32
    assert _ham(lib, prefix, b"A" * 33, b"A" * 32 + b"T") == 1
    assert _ham(lib, prefix, b"C" + b"A" * 32, b"A" * 32 + b"T") == 2
    assert _ham(lib, prefix, b"A" * 31, b"A" * 15 + b"G" + b"A" * 15) == 1
    assert _ham(lib, prefix, b"A", b"A") == 0
    assert _ham(lib, prefix, b"A", b"T") == 1
    assert _ham(lib, prefix, b"", b"") == 0


@pytest.mark.parametrize("prefix,lib", LIBS)
def test_has_repeat_known_answers(prefix, lib):
    # repeat_test.cpp:163-212
    assert _rep_kmers(lib, prefix, [b"ACGT", b"TGCA", b"ACGT", b"GGCC"], 0)
    assert not _rep_kmers(lib, prefix, [b"ACGT", b"TGCA", b"GGCC", b"AATT"], 0)
    assert not _rep_kmers(lib, prefix, [], 0)
    assert not _rep_kmers(lib, prefix, [b"ACGT"], 0)
    assert _rep_kmers(lib, prefix, [b"ACGT", b"TGCA", b"ACGA"], 1)
    assert not _rep_kmers(lib, prefix, [b"ACGT", b"TGCA", b"ACGA"], 0)
    assert not _rep_kmers(lib, prefix, [b"AAAA", b"CCCC", b"GGGG", b"TTTT"], 1)


@pytest.mark.parametrize("prefix,lib", LIBS)
def test_hamming_random_dna(prefix, lib):
    # repeat_test.cpp:52-71 and :245-280: identical → 0, against a scalar count otherwise
    rng = np.random.default_rng(0x5EED)
    for n in (1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 5000):
        a = np.frombuffer(repeat_lib.random_window(rng, n), dtype=np.uint8)
        b = np.frombuffer(repeat_lib.random_window(rng, n), dtype=np.uint8)
        assert _ham(lib, prefix, a.tobytes(), a.tobytes()) == 0
        assert _ham(lib, prefix, a.tobytes(), b.tobytes()) == int((a != b).sum())


def test_sliding_semantics():
    # base::SlidingView (sliding.h:17-34): k-mers at every offset; fewer than two → never a repeat
    assert ORC.orc_has_repeat(b"ACGT", 4, 4, 0) == 0
    assert ORC.orc_has_repeat(b"ACG", 3, 4, 2) == 0
    assert ORC.orc_has_repeat(b"", 0, 4, 2) == 0
    assert ORC.orc_has_repeat(b"AAAAA", 5, 4, 0) == 1            # AAAA at 0 and 1
    assert ORC.orc_has_repeat(b"ACGTACGT", 8, 4, 0) == 1
    assert ORC.orc_has_repeat(b"ACGTACGA", 8, 4, 0) == 0
    assert ORC.orc_has_repeat(b"ACGTACGA", 8, 4, 1) == 1
    assert ORC.orc_has_repeat(b"ACGTacgt", 8, 4, 0) == 0         # raw bytes: case matters, as in the reference


@pytest.mark.skipif(REF is None, reason="oracle/_ref/liblancet_ref_repeat.so not built (needs /root/reference)")
def test_oracle_matches_live_reference_on_windows():
    jobs = repeat_lib.window_jobs(seed=20260801, n_windows=48, k_values=(5, 11, 13, 21, 33, 64, 101), lengths=(120, 400, 700))
    seen = set()
    for seq, k, mm in jobs:
        for m in (mm, 0, 1, 3):
            got = ORC.orc_has_repeat(seq, len(seq), k, m)
            want = REF.ref_has_repeat(seq, len(seq), k, m)
            assert got == want, (len(seq), k, m)
            seen.add(got)
    assert seen == {0, 1}


def test_min_distance_consistency():
    rng = np.random.default_rng(7)
    for _ in range(40):
        n, k = int(rng.integers(20, 200)), int(rng.integers(3, 20))
        seq = repeat_lib.random_window(rng, n, b"ACGT" if rng.integers(0, 2) else b"AC")
        dmin = ORC.orc_min_kmer_distance(seq, n, k)
        for mm in range(0, 5):
            assert ORC.orc_has_repeat(seq, n, k, mm) == int(dmin >= 0 and dmin <= mm)
