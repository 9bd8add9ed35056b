"""Loaders and generators shared by the repeat-detection tests."""
import ctypes as C
import os

import numpy as np

from oracle_lib import ORACLE_DIR, load_oracle

REF_REPEAT_SO = os.path.join(ORACLE_DIR, "_ref", "liblancet_ref_repeat.so")


def _bind(lib, prefix):
    f = getattr(lib, prefix + "_hamming_dist")
    f.argtypes, f.restype = [C.c_char_p, C.c_char_p, C.c_int64], C.c_uint64
    f = getattr(lib, prefix + "_has_repeat")
    f.argtypes, f.restype = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64], C.c_int
    f = getattr(lib, prefix + "_has_repeat_kmers")
    f.argtypes, f.restype = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64], C.c_int
    return lib


def oracle():
    lib = _bind(load_oracle(), "orc")
    lib.orc_min_kmer_distance.argtypes, lib.orc_min_kmer_distance.restype = [C.c_char_p, C.c_int64, C.c_int64], C.c_int64
    return lib


def reference():
    """The reference's own base/repeat.cpp compiled unmodified into oracle/_ref (None when not built)."""
    return _bind(C.CDLL(REF_REPEAT_SO), "ref") if os.path.exists(REF_REPEAT_SO) else None


def random_window(rng, length, alphabet=b"ACGT"):
    return bytes(np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), length)])


def plant_repeat(rng, seq, k, mismatches, gap=None):
    """Copy one k-mer of seq to another offset with exactly `mismatches` substituted bases."""
    s = bytearray(seq)
    n = len(s)
    if n < k + 1:
        return bytes(s)
    src = int(rng.integers(0, max(1, n - 2 * k))) if gap is None else 0
    dst = int(rng.integers(min(src + k, n - k), n - k + 1)) if gap is None else gap  # clear of the source when it fits
    kmer = bytearray(s[src:src + k])
    for p in rng.choice(k, size=mismatches, replace=False):
        kmer[p] = ord("ACGT"[("ACGT".index(chr(kmer[p])) + 1 + int(rng.integers(0, 3))) % 4])
    s[dst:dst + k] = kmer
    return bytes(s)


def window_jobs(seed, n_windows, k_values=(13, 19, 25, 31, 61, 127), lengths=(600, 1000, 1400)):
    """(sequence, k, max_mismatches) jobs shaped like the callers': every window asked at several k with 2
    mismatches allowed (Graph's k-loop) and once exactly at the largest k (ShouldSkipWindow).  A third of
    the windows carry a planted approximate repeat, some a low-complexity stretch."""
    rng = np.random.default_rng(seed)
    jobs = []
    for w in range(n_windows):
        n = int(lengths[w % len(lengths)]) + int(rng.integers(0, 50))
        seq = random_window(rng, n)
        kind = w % 6
        if kind == 1:
            seq = plant_repeat(rng, seq, int(rng.choice(k_values)), int(rng.integers(0, 4)))
        elif kind == 3:
            unit = random_window(rng, int(rng.integers(1, 7)))
            at = int(rng.integers(0, max(1, n - 200)))
            reps = int(rng.integers(4, 40))
            seq = (seq[:at] + unit * reps + seq[at:])[:n]
        elif kind == 5:
            seq = plant_repeat(rng, seq, int(rng.choice(k_values)), 3)
        for k in k_values:
            jobs.append((seq, k, 2))
        jobs.append((seq, max(k_values), 0))
    return jobs
