#!/usr/bin/env python3
"""Generate tests/golden/evidence_golden.json with the REFERENCE's own
VariantSupport::AddEvidence (src/lancet/caller/variant_support.cpp:23-67 compiled unmodified into
oracle/_ref by `make -C oracle ref`).  Run in the build container only."""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FIELDS = [("isize", np.int64), ("start", np.int64), ("aln", np.float64), ("fold", np.float64), ("hash", np.uint32),
          ("ref_nm", np.uint32), ("own_nm", np.uint32), ("hap_id", np.uint32), ("allele", np.uint8), ("rev", np.uint8),
          ("bq", np.uint8), ("mapq", np.uint8), ("softclip", np.uint8), ("proper", np.uint8)]


def call_dump(fn, st):
    arrs = [np.ascontiguousarray(st[k], dtype=dt) for k, dt in FIELDS]
    buf = C.create_string_buffer(1 << 20)
    fn.argtypes = [C.c_int] + [C.c_void_p] * 14 + [C.c_char_p, C.c_longlong]
    fn.restype = C.c_int
    n = fn(len(arrs[0]), *[a.ctypes.data for a in arrs], buf, len(buf))
    assert n >= 0
    return buf.value.decode()


def random_stream(rng):
    n = int(rng.integers(1, 60))
    names = rng.integers(0, max(2, n // 2), size=n)  # repeated name hashes (mates) exercise the dedup
    return dict(
        isize=rng.integers(-600, 600, n) * (rng.random(n) < 0.9), start=rng.integers(1000, 5000, n),
        aln=np.round(rng.normal(120, 30, n), 3), fold=rng.random(n) / 2, hash=(names * 2654435761) % (1 << 32),
        ref_nm=rng.integers(0, 9, n), own_nm=rng.integers(0, 5, n), hap_id=rng.integers(0, 4, n),
        allele=rng.integers(0, 3, n), rev=rng.integers(0, 2, n), bq=rng.integers(0, 42, n), mapq=rng.integers(0, 61, n),
        softclip=rng.integers(0, 2, n), proper=rng.integers(0, 2, n))


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblancet_ref_scoring.so"))
    lib.ref_add_evidence_dump.restype = C.c_int
    rng = np.random.default_rng(77)
    out = []
    for _ in range(60):
        st = random_stream(rng)
        out.append({"stream": {k: np.asarray(v).tolist() for k, v in st.items()}, "dump": call_dump(lib.ref_add_evidence_dump, st)})
    json.dump({"source": "reference variant_support.cpp AddEvidence compiled unmodified", "cases": out},
              open(os.path.join(HERE, "evidence_golden.json"), "w"), separators=(",", ":"))
    print("wrote", len(out))


if __name__ == "__main__":
    main()
