#!/usr/bin/env python3
"""Generate tests/golden/scoring_golden.json from the REFERENCE's own scoring
code (oracle/_ref/liblancet_ref_scoring.so = /root/reference/src/lancet/caller/
{local_scorer,combined_scorer}.cpp + hts/phred_quality.cpp + hts/cigar_utils.h
compiled unmodified; recipe: oracle/Makefile target `ref`).

Run in the build container only (needs /root/reference):
    make -C oracle ref && python tests/golden/gen_scoring_golden.py
The JSON travels with the repo; /root/reference does not exist on the GPU box.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblancet_ref_scoring.so")
OPS = "MIDNSHP=XB"


def load_ref():
    lib = C.CDLL(REF_SO)
    lib.ref_phred_err.argtypes = [C.c_uint32]
    lib.ref_phred_err.restype = C.c_double
    lib.ref_encode.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    lib.ref_edit_distance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.ref_edit_distance.restype = C.c_uint32
    lib.ref_refpos_to_qpos.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    lib.ref_refpos_to_qpos.restype = C.c_uint64
    lib.ref_softclip_penalty.argtypes = [C.c_void_p, C.c_int]
    lib.ref_softclip_penalty.restype = C.c_double
    lib.ref_local_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ref_score_read_at_variant.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int,
                                              C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int32,
                                              C.c_int32, C.c_int, C.c_void_p, C.c_void_p]
    lib.ref_hap_edit_distance.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_int]
    lib.ref_hap_edit_distance.restype = C.c_uint32
    return lib


def random_case(rng):
    """random read/hap/cigar that is self-consistent (cigar consumes the read fully)"""
    qlen = int(rng.integers(20, 160))
    ops = []
    q_left = qlen
    if rng.random() < 0.3:
        s = int(rng.integers(1, 10)); ops.append((4, s)); q_left -= s
    tail_s = int(rng.integers(1, 10)) if rng.random() < 0.3 else 0
    q_left -= tail_s
    tlen = 0
    first = True
    while q_left > 0:
        r = rng.random()
        if first or r < 0.6:
            l = int(rng.integers(1, max(2, q_left + 1))); l = min(l, q_left)
            ops.append((0, l)); q_left -= l; tlen += l
        elif r < 0.8:
            l = min(int(rng.integers(1, 12)), q_left); ops.append((1, l)); q_left -= l
        else:
            l = int(rng.integers(1, 30)); ops.append((2, l)); tlen += l
        first = False
    if tail_s:
        ops.append((4, tail_s))
    rs = int(rng.integers(0, 50))
    hap_n = rs + tlen + int(rng.integers(0, 50))
    alpha = np.array([0, 1, 2, 3, 4], dtype=np.uint8)
    pq = [0.245, 0.245, 0.245, 0.245, 0.02]
    hap = rng.choice(alpha, size=hap_n, p=pq)
    read = rng.choice(alpha, size=qlen, p=pq)
    # make M blocks mostly matching
    qp, tp = 0, rs
    for op, l in ops:
        if op == 0:
            m = rng.random(l) < 0.92
            read[qp:qp + l][m] = hap[tp:tp + l][m]
            qp += l; tp += l
        elif op in (1, 4):
            qp += l
        elif op == 2:
            tp += l
    quals = rng.choice(np.array([2, 12, 23, 37, 41, 93, 255], dtype=np.uint8), size=qlen)
    vstart = int(rng.integers(max(0, rs - 5), rs + tlen + 5))
    vlen = int(rng.integers(0, 40))
    return dict(cigar=[(l << 4) | op for op, l in ops], rs=rs, re=rs + tlen, hap=hap.tolist(), read=read.tolist(),
                quals=quals.tolist(), var_start=vstart, var_len=vlen, score=int(rng.integers(30, 160)),
                hap_idx=int(rng.integers(0, 4)), allele=int(rng.integers(0, 3)))


def main():
    lib = load_ref()
    rng = np.random.default_rng(20261017)
    cases = []
    for _ in range(300):
        c = random_case(rng)
        cig = np.asarray(c["cigar"], dtype=np.uint32)
        hap = np.asarray(c["hap"], dtype=np.uint8)
        read = np.asarray(c["read"], dtype=np.uint8)
        quals = np.asarray(c["quals"], dtype=np.uint8)
        tgt = hap[c["rs"]:c["re"]].copy()
        of = np.zeros(4); oi = np.zeros(5, dtype=np.int64)
        lib.ref_score_read_at_variant(cig.ctypes.data, cig.size, c["score"], c["rs"], c["re"], c["hap_idx"],
                                      hap.ctypes.data, hap.size, read.ctypes.data, read.size, quals.ctypes.data,
                                      c["var_start"], c["var_len"], c["allele"], of.ctypes.data, oi.ctypes.data)
        o3 = np.zeros(3); bq = C.c_uint8(0)
        lib.ref_local_score(cig.ctypes.data, cig.size, read.ctypes.data, read.size, tgt.ctypes.data, tgt.size,
                            quals.ctypes.data, quals.size, c["rs"], c["var_start"], c["var_len"], o3.ctypes.data,
                            C.byref(bq))
        c["expect"] = dict(
            local_score=of[0].hex(), local_identity=of[1].hex(), folded=of[2].hex(), combined=of[3].hex(),
            global_score=int(oi[0]), own_nm=int(oi[1]), hap_id=int(oi[2]), allele=int(oi[3]), base_qual=int(oi[4]),
            pbq=o3[0].hex(), raw=o3[1].hex(), identity=o3[2].hex(), min_bq=int(bq.value),
            nm=int(lib.ref_edit_distance(cig.ctypes.data, cig.size, read.ctypes.data, read.size, tgt.ctypes.data, tgt.size)),
            sc_pen=float(lib.ref_softclip_penalty(cig.ctypes.data, cig.size)),
            qpos=[int(lib.ref_refpos_to_qpos(cig.ctypes.data, cig.size, rp)) for rp in (0, 1, 5, 17, 60, 10000)],
            hap_nm_match=int(lib.ref_hap_edit_distance(cig.ctypes.data, cig.size, c["rs"], c["re"], 0, 0, hap.ctypes.data, hap.size, read.ctypes.data, read.size)),
            hap_nm_miss=int(lib.ref_hap_edit_distance(cig.ctypes.data, cig.size, c["rs"], c["re"], 1, 0, hap.ctypes.data, hap.size, read.ctypes.data, read.size)),
        )
        cases.append(c)
    phred = [float(lib.ref_phred_err(q)).hex() for q in range(256)]
    enc = np.zeros(256, dtype=np.uint8)
    allb = bytes(range(256))
    lib.ref_encode(allb, 256, enc.ctypes.data)
    out = dict(source="reference local_scorer.cpp/combined_scorer.cpp/phred_quality.cpp/cigar_utils.h compiled unmodified",
               phred=phred, encode=enc.tolist(), cases=cases)
    with open(os.path.join(HERE, "scoring_golden.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
