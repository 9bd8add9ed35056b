"""Synthetic realignment workloads: deterministic groups of
haplotypes / reads / variants shaped like what `Genotyper::Genotype` receives
(reference: src/lancet/caller/genotyper.cpp:224-235).

microbench group: hap 0 random ACGT; haps 1..P-1 = hap 0 + one spiked variant
(70 % SNV, 20 % indel 1-20 bp, 10 % indel 21-300 bp); R reads of length L sampled
uniformly from the P haplotypes (5 % overhang an end by <= 30 bp), 0.2 %
substitution errors, 0.01 % indel errors, qualities from {Q12 3 %, Q23 7 %,
Q37 90 %}; V = P-1 variants with per-haplotype bounds in VCF style (anchor base
for indels), as VariantExtractor would hand them to ExtractHapBounds
(genotyper.cpp:329-352).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .abi import Group

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_QV = np.array([12, 23, 37], dtype=np.uint8)
_QP = np.array([0.03, 0.07, 0.90])


def _rand_bases(rng: np.random.Generator, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n)]


def _spike_variant(rng: np.random.Generator, ref: np.ndarray) -> Tuple[np.ndarray, tuple, tuple]:
    """returns (alt_hap, ref_bounds(start,len), alt_bounds(start,len))"""
    n = ref.size
    lo, hi = 50, max(51, n - 50)
    u = rng.random()
    pos = int(rng.integers(lo, hi))
    if u < 0.70:  # SNV
        alt = ref.copy()
        alt[pos] = _ACGT[(int(np.where(_ACGT == ref[pos])[0][0]) + int(rng.integers(1, 4))) % 4]
        return alt, (pos, 1), (pos, 1)
    ln = int(rng.integers(1, 21)) if u < 0.90 else int(rng.integers(21, 301))
    if rng.random() < 0.5:  # insertion after anchor base `pos`
        ins = _rand_bases(rng, ln)
        alt = np.concatenate([ref[:pos + 1], ins, ref[pos + 1:]])
        return alt, (pos, 1), (pos, 1 + ln)
    ln = min(ln, n - pos - 30)  # deletion of ln bases after the anchor
    ln = max(ln, 1)
    alt = np.concatenate([ref[:pos + 1], ref[pos + 1 + ln:]])
    return alt, (pos, 1 + ln), (pos, 1)


def make_group(rng: np.random.Generator, read_len: int = 150, hap_len: int = 1000, n_haps: int = 4,
               n_reads: int = 512, name_prefix: str = "r", sub_err: float = 0.002, indel_err: float = 0.0001,
               overhang_frac: float = 0.05, n_frac: float = 0.0) -> Group:
    ref = _rand_bases(rng, hap_len)
    haps = [ref]
    variants: List[List[tuple]] = []
    for h in range(1, n_haps):
        alt, rb, ab = _spike_variant(rng, ref)
        haps.append(alt)
        row = [(-1, 0, -1)] * n_haps
        row[0] = (rb[0], rb[1], 0)
        row[h] = (ab[0], ab[1], 1)
        variants.append(row)
    reads, quals, names = [], [], []
    src = rng.integers(0, n_haps, size=n_reads)
    over = rng.random(n_reads) < overhang_frac
    for i in range(n_reads):
        hp = haps[int(src[i])]
        hl = hp.size
        if over[i]:
            oh = int(rng.integers(1, 31))
            if rng.random() < 0.5:
                st = -oh
            else:
                st = hl - read_len + oh
        else:
            st = int(rng.integers(0, max(1, hl - read_len + 1)))
        idx = np.arange(st, st + read_len)
        inside = (idx >= 0) & (idx < hl)
        rd = _rand_bases(rng, read_len)
        rd[inside] = hp[idx[inside]]
        errs = rng.random(read_len) < sub_err
        if errs.any():
            rd[errs] = _rand_bases(rng, int(errs.sum()))
        if indel_err > 0 and rng.random() < indel_err * read_len:
            p = int(rng.integers(5, read_len - 5))
            if rng.random() < 0.5:
                rd = np.concatenate([rd[:p], _rand_bases(rng, 1), rd[p:-1]])
            else:
                rd = np.concatenate([rd[:p], rd[p + 1:], _rand_bases(rng, 1)])
        if n_frac > 0:
            nm = rng.random(read_len) < n_frac
            rd[nm] = ord("N")
        q = _QV[rng.choice(3, size=read_len, p=_QP)]
        reads.append(rd.tobytes())
        quals.append(q.tobytes())
        names.append(f"{name_prefix}{i}")
    return Group(haps=[h.tobytes() for h in haps], reads=reads, quals=quals, names=names, variants=variants)


def make_groups(seed: int, n_groups: int, **kw) -> List[Group]:
    rng = np.random.default_rng(seed)
    return [make_group(rng, name_prefix=f"g{g}r", **kw) for g in range(n_groups)]
