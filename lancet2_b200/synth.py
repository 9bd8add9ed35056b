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

from typing import List, Sequence, Tuple

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


# --------------------------------------------------------------------------------------
# BASELINE.json config 2: "synthetic 30x/30x tumor-normal 2x150bp reads over a 1 Mb
# synthetic reference with spiked SNVs/InDels".  Graph assembly / SPOA stay host code in the
# reference and are out of scope here, so this generator emits what they would hand to
# Genotyper::Genotype for every 1000 bp window (step 800, core/window_builder.h:19-38) that
# contains a variant: the window's REF haplotype, one ALT haplotype per spiked variant (plus
# the all-variants haplotype when a window holds several), every tumor/normal read that
# overlaps the window in ReadCollector order (normal before tumor, then qname), and the
# per-haplotype variant bounds.
# --------------------------------------------------------------------------------------
def make_region_groups(seed: int = 42, ref_len: int = 1_000_000, cov_normal: float = 30.0, cov_tumor: float = 30.0,
                       read_len: int = 150, window: int = 1000, step: int = 800, var_every: int = 2000,
                       sub_err: float = 0.002, max_groups: int = 0, samples=None) -> List[Group]:
    """samples: optional list of (read-name tag, coverage, is_tumor) in ReadCollector order (tag kind —
    normal before tumor —, then sample name: core/read_collector.cpp:42-54).  Default: one normal
    ("n") and one tumor ("t") at cov_normal / cov_tumor, i.e. BASELINE configs[1]/[2].  Every group
    carries `sample`, the index into `samples` of each read."""
    rng = np.random.default_rng(seed)
    if samples is None:
        samples = [("n", cov_normal, False), ("t", cov_tumor, True)]
    sample_spec = list(samples)
    ref = _rand_bases(rng, ref_len)
    # ---- spiked variants (sorted, spaced so that they never overlap) ----
    n_var = max(1, ref_len // var_every)
    pos = np.sort(rng.choice(np.arange(400, ref_len - 800, 400), size=min(n_var, (ref_len - 1200) // 400), replace=False))
    pos = pos + rng.integers(0, 50, size=pos.size)
    variants = []  # (pos, ref_allele bytes, alt_allele bytes, vaf_normal, vaf_tumor)
    for p in pos.tolist():
        u = rng.random()
        if u < 0.70:
            alt_b = _ACGT[(int(np.where(_ACGT == ref[p])[0][0]) + int(rng.integers(1, 4))) % 4]
            ra, aa = ref[p:p + 1].tobytes(), bytes([alt_b])
        else:
            ln = int(rng.integers(1, 21)) if u < 0.90 else int(rng.integers(21, 301))
            if rng.random() < 0.5:
                ra, aa = ref[p:p + 1].tobytes(), ref[p:p + 1].tobytes() + _rand_bases(rng, ln).tobytes()
            else:
                ra, aa = ref[p:p + 1 + ln].tobytes(), ref[p:p + 1].tobytes()
        c = rng.random()
        if c < 0.5:
            vn, vt = 0.0, float(rng.choice([0.05, 0.1, 0.25, 0.5]))
        elif c < 0.8:
            vn = vt = 0.5
        else:
            vn = vt = 1.0
        variants.append((p, ra, aa, vn, vt))
    vpos = np.asarray([v[0] for v in variants], dtype=np.int64)

    # ---- reads ----
    def sample_reads(cov, is_tumor, tag):
        n_frag = int(cov * ref_len / (2 * read_len))
        ins = np.clip(rng.normal(400, 50, n_frag).astype(np.int64), 2 * read_len // 2 + 20, 800)
        fs = rng.integers(0, ref_len - 900, size=n_frag)
        starts = np.concatenate([fs, fs + ins - read_len])
        frag_id = np.concatenate([np.arange(n_frag), np.arange(n_frag)])
        idx = starts[:, None] + np.arange(read_len)[None, :]
        seqs = ref[idx]
        # reads spanning a variant: rebuild from the mutated sequence with probability VAF
        lo = np.searchsorted(vpos, starts - 301, side="left")
        hi = np.searchsorted(vpos, starts + read_len, side="left")
        for i in np.nonzero(hi > lo)[0].tolist():
            s = int(starts[i])
            out, cur = [], s
            need = read_len
            for k in range(int(lo[i]), int(hi[i])):
                p, ra, aa, vn, vt = variants[k]
                vaf = vt if is_tumor else vn
                if p + len(ra) <= cur or rng.random() >= vaf:
                    continue
                if p < cur:  # read starts inside the ref allele: skip the variant
                    continue
                out.append(ref[cur:p].tobytes())
                out.append(aa)
                cur = p + len(ra)
                if sum(map(len, out)) >= need:
                    break
            got = b"".join(out)
            if len(got) < need:
                got += ref[cur:cur + need - len(got)].tobytes()
            seqs[i] = np.frombuffer(got[:need], dtype=np.uint8)
        err = rng.random(seqs.shape) < sub_err
        seqs[err] = _ACGT[rng.integers(0, 4, size=int(err.sum()))]
        quals = _QV[rng.choice(3, size=seqs.shape, p=_QP)]
        names = np.asarray([f"{tag}{f:07d}" for f in frag_id.tolist()])
        return starts, seqs, quals, names

    samples = [sample_reads(cov, is_tumor, tag) for tag, cov, is_tumor in sample_spec]
    order = [np.argsort(s[0], kind="stable") for s in samples]

    groups: List[Group] = []
    for w0 in range(0, ref_len - window + 1, step):
        k0, k1 = np.searchsorted(vpos, w0 + 50), np.searchsorted(vpos, w0 + window - 50 - 301)
        if k1 <= k0:
            continue
        wv = [variants[k] for k in range(k0, k1)]
        wv = [v for v in wv if v[0] + len(v[1]) < w0 + window - 20]
        if not wv:
            continue
        ref_hap = ref[w0:w0 + window].tobytes()
        hap_sets = [[i] for i in range(len(wv))]
        if len(wv) > 1:
            hap_sets.append(list(range(len(wv))))
        haps = [ref_hap]
        shifts = []  # per alt hap: dict variant index -> start on that hap
        for hs in hap_sets:
            out, cur, sh, starts_on_hap = [], 0, 0, {}
            for vi in hs:
                p, ra, aa, _, _ = wv[vi]
                lp = p - w0
                out.append(ref_hap[cur:lp])
                starts_on_hap[vi] = lp + sh
                out.append(aa)
                cur = lp + len(ra)
                sh += len(aa) - len(ra)
            out.append(ref_hap[cur:])
            haps.append(b"".join(out))
            shifts.append(starts_on_hap)
        rows = []
        for vi, (p, ra, aa, _, _) in enumerate(wv):
            row = [(-1, 0, -1)] * len(haps)
            row[0] = (p - w0, len(ra), 0)
            for hi_, soh in enumerate(shifts):
                if vi in soh:
                    row[hi_ + 1] = (soh[vi], len(aa), 1)
            rows.append(row)
        reads, quals, names, sample_of = [], [], [], []
        for si, ((starts, seqs, qv, nm), od) in enumerate(zip(samples, order)):
            ss = starts[od]
            a, b = np.searchsorted(ss, w0 - read_len + 1), np.searchsorted(ss, w0 + window)
            sel = od[a:b]
            sel = sel[np.argsort(nm[sel], kind="stable")]
            for i in sel.tolist():
                reads.append(seqs[i].tobytes())
                quals.append(qv[i].tobytes())
                names.append(str(nm[i]))
                sample_of.append(si)
        groups.append(Group(haps=haps, reads=reads, quals=quals, names=names, variants=rows, sample=sample_of))
        if max_groups and len(groups) >= max_groups:
            break
    return groups


# --------------------------------------------------------------------------------------
# BASELINE.json configs[2] / configs[3], the multi-GPU workloads.  A ~50 Mb (cfg3) or 10 Mb (cfg4)
# reference is laid out as independent 1 Mb tiles (contigs), tile t generated from seed
# 1000 * seed + t, so that every rank of a sharded run generates exactly the tiles it owns and
# the whole workload is the same whatever the number of ranks.
#   cfg3: 60x tumor / 40x normal, 50 tiles
#   cfg4: four samples at 30x each (one normal, three tumors: colored-graph multi-sample calling), 10 tiles
# --------------------------------------------------------------------------------------
TILED = {
    "cfg3": dict(tiles=50, samples=[("n", 40.0, False), ("t", 60.0, True)],
                 sample_names=["normal", "tumor"]),
    "cfg4": dict(tiles=10, samples=[("a", 30.0, False), ("b", 30.0, True), ("c", 30.0, True), ("d", 30.0, True)],
                 sample_names=["S1_normal", "S2_tumor", "S3_tumor", "S4_tumor"]),
}


def tile_cost(name: str, seed: int, tile: int) -> int:
    """work estimate of one tile for the shard partition (lancet2_b200/dispatch.py): the tiles of one
    workload are statistically identical (same length, coverage and variant density)"""
    spec = TILED[name]
    return int(sum(c for _, c, _ in spec["samples"]) * 1_000_000)


def make_tile_groups(name: str, seed: int, tile: int, ref_len: int = 1_000_000) -> List[Group]:
    spec = TILED[name]
    return make_region_groups(1000 * seed + tile, ref_len=ref_len, samples=spec["samples"])


def _tile_job(args):
    return make_tile_groups(*args)


def make_tiled_groups(name: str, seed: int, tiles: Sequence[int], ref_len: int = 1_000_000, procs: int = 1, per_tile: bool = False):
    """the groups of the given tiles, in tile order (per_tile: one list per tile); `procs` > 1 generates tiles in
    worker processes"""
    jobs = [(name, seed, int(t), ref_len) for t in tiles]
    if procs > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
            parts = pool.map(_tile_job, jobs, chunksize=1)
    else:
        parts = [_tile_job(j) for j in jobs]
    return parts if per_tile else [g for part in parts for g in part]
