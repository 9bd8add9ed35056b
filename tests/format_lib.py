"""TEST INFRASTRUCTURE for the FORMAT-math row: loads the g++ build of the
device core (tests/hostemu/format_emu.cpp), the REFERENCE's own VariantSupport when
oracle/_ref is present, random evidence streams and the field-by-field comparison."""
import ctypes as C
import os
import subprocess

import numpy as np

from lancet2_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_SRC = os.path.join(HERE, "hostemu", "format_emu.cpp")
EMU_SO = os.path.join(HERE, "hostemu", "libformat_emu.so")
CORE = os.path.join(ROOT, "lancet2_b200", "csrc", "lgr_format.cuh")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblancet_ref_scoring.so")
REL_TOL = 1e-9   # f64 metrics: warp-tree sums + another libm vs the reference's in-order sums
ABS_TOL = 1e-9
_emu = None
_ref = None


def load_emu():
    global _emu
    if _emu is None:
        deps = [EMU_SRC, CORE, os.path.join(ROOT, "include", "lancet_gpu_realign.h")]
        if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                                   "-Wno-unknown-pragmas", "-o", EMU_SO, EMU_SRC])
        _emu = C.CDLL(EMU_SO)
        _emu.emu_format_metrics.argtypes = [C.POINTER(abi.LgrEvidenceIn), C.c_void_p]
        _emu.emu_format_metrics.restype = C.c_int
        _emu.emu_format_metrics_ex.argtypes = [C.POINTER(abi.LgrEvidenceIn), C.c_void_p, C.c_int]
        _emu.emu_format_metrics_ex.restype = C.c_int
    return _emu


_emu_sort = None
EMU_SORT_SO = os.path.join(HERE, "hostemu", "libformat_emu_sort.so")


def load_emu_sort():
    """the same core compiled with -DLGR_FMT_SORT (shared-memory sort instead of the
    O(n^2) scans; checked on the CPU only, not in the default device build)"""
    global _emu_sort
    if _emu_sort is None:
        deps = [EMU_SRC, CORE, os.path.join(ROOT, "include", "lancet_gpu_realign.h")]
        if not os.path.exists(EMU_SORT_SO) or os.path.getmtime(EMU_SORT_SO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DLGR_FMT_SORT",
                                   "-Wno-unknown-pragmas", "-o", EMU_SORT_SO, EMU_SRC])
        _emu_sort = C.CDLL(EMU_SORT_SO)
        _emu_sort.emu_format_metrics_ex.argtypes = [C.POINTER(abi.LgrEvidenceIn), C.c_void_p, C.c_int]
        _emu_sort.emu_format_metrics_ex.restype = C.c_int
    return _emu_sort


def emu_format_sort(supports):
    batch = supports if isinstance(supports, abi.EvidenceBatch) else abi.EvidenceBatch(supports)
    out = np.zeros(batch.n_supports, dtype=abi.FORMAT_DTYPE)
    st = batch.c_struct()
    rc = load_emu_sort().emu_format_metrics_ex(C.byref(st), out.ctypes.data, 1)
    return rc, out


def emu_format(supports, split_tasks=True):
    """split_tasks: one call of the core per (support, task) — what k_fmt_metrics launches; False: all
    tasks of a support in one call."""
    batch = supports if isinstance(supports, abi.EvidenceBatch) else abi.EvidenceBatch(supports)
    out = np.zeros(batch.n_supports, dtype=abi.FORMAT_DTYPE)
    st = batch.c_struct()
    rc = load_emu().emu_format_metrics_ex(C.byref(st), out.ctypes.data, 1 if split_tasks else 0)
    return rc, out


def have_ref():
    return os.path.exists(REF_SO)


def ref_format(supports):
    """The reference's VariantSupport (compiled unmodified into oracle/_ref) on every support."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
        _ref.ref_support_metrics.argtypes = [C.c_int] + [C.c_void_p] * 12 + [C.c_int] * 3 + [C.c_void_p]
        _ref.ref_support_metrics.restype = C.c_int
    out = np.zeros(len(supports), dtype=abi.FORMAT_DTYPE)
    for s, sup in enumerate(supports):
        cols = [np.ascontiguousarray(sup[name], dtype=dt) for name, dt in abi.EVIDENCE_FIELDS]
        rc = _ref.ref_support_metrics(len(cols[0]), *[c.ctypes.data for c in cols], int(sup["n_alleles"]),
                                      int(sup.get("variant_len", 0)), int(sup.get("total_haps", 2)),
                                      out[s:s + 1].ctypes.data)
        assert rc == 0
    return out


def ref_format_batch(batch):
    """one call of the reference's VariantSupport over a whole EvidenceBatch (timing baseline)"""
    ref_format([])  # loads the library
    _ref.ref_support_metrics_batch.argtypes = [C.POINTER(abi.LgrEvidenceIn), C.c_void_p]
    _ref.ref_support_metrics_batch.restype = C.c_int
    out = np.zeros(batch.n_supports, dtype=abi.FORMAT_DTYPE)
    st = batch.c_struct()
    rc = _ref.ref_support_metrics_batch(C.byref(st), out.ctypes.data)
    assert rc == 0
    return out


def random_support(rng, n=None, n_alleles=None, dup_frac=0.3):
    n = int(rng.integers(0, 80)) if n is None else n
    k = int(rng.integers(1, 5)) if n_alleles is None else n_alleles
    names = rng.integers(0, max(2, int(n * (1 - dup_frac)) + 1), size=n)  # repeated name hashes (mates) → dedup
    quals = np.array([2, 12, 23, 37, 40, 0, 93], dtype=np.uint8)
    p_alt = rng.choice([0.0, 0.05, 0.3, 0.5, 1.0])
    allele = np.where(rng.random(n) < p_alt, rng.integers(1, max(2, k), n), 0) if k > 1 else np.zeros(n, int)
    fold = rng.random(n) / 2
    if n and rng.random() < 0.5:
        fold = np.round(fold, 1)  # ties in the f64 Mann-Whitney
    return dict(
        insert_size=rng.integers(-600, 600, n) * (rng.random(n) < 0.9), aln_start=rng.integers(-50, 3000, n),
        aln_score=np.round(rng.normal(120, 30, n), 3), folded_pos=fold, rname_hash=(names * 2654435761) % (1 << 32),
        ref_nm=rng.integers(0, 60, n), own_hap_nm=rng.integers(0, 5, n), hap_id=rng.integers(0, 4, n),
        allele=np.minimum(allele, k - 1), flags=rng.integers(0, 8, n),
        base_qual=quals[rng.integers(0, len(quals), n)] if rng.random() < 0.7 else rng.integers(0, 60, n),
        map_qual=rng.choice([0, 20, 40, 60], n), n_alleles=k, variant_len=int(rng.integers(0, 50)),
        total_haps=int(rng.integers(1, 6)))


def simple_support(rows, n_alleles=2, variant_len=0, total_haps=2):
    """rows of (allele, aln_start, own_hap_nm, hap_id, rname_hash) with the fixed other fields of the
    reference's MakeEvidence helper (tests/caller/variant_support_metrics_test.cpp:12-30)."""
    n = len(rows)
    r = np.asarray(rows, dtype=np.int64).reshape(n, 5)
    return dict(insert_size=np.full(n, 300), aln_start=r[:, 1], aln_score=np.full(n, 100.0),
                folded_pos=np.full(n, 0.25), rname_hash=r[:, 4], ref_nm=np.zeros(n, int), own_hap_nm=r[:, 2],
                hap_id=r[:, 3], allele=r[:, 0], flags=np.full(n, abi.LGR_EV_PROPER_PAIR), base_qual=np.full(n, 30),
                map_qual=np.full(n, 60), n_alleles=n_alleles, variant_len=variant_len, total_haps=total_haps)


EXACT_FIELDS = ("fwd", "rev", "soft_clip", "n_alleles", "valid", "n_kept", "pl", "gq")
MW_FIELDS = ("mqcd", "rpcd", "bqcd")           # integer rank statistics → same bits
F64_FIELDS = ("raw_pbq", "rms_mq", "mean_aln", "cmlod", "sb", "sca", "fld", "asmd", "fsse", "ahdd", "hse")


def compare_format(want, got, exact_pl=True, label="", mw_exact=True):
    """Field-by-field differences between two FORMAT_DTYPE arrays (list of strings, empty = equal)."""
    errs = []
    for f in EXACT_FIELDS:
        if f in ("pl", "gq") and not exact_pl:
            bad = np.argwhere(np.abs(want[f].astype(np.int64) - got[f].astype(np.int64)) > 1)
        else:
            bad = np.argwhere(want[f] != got[f])
        for idx in bad[:5]:
            errs.append(f"{label}{f}{tuple(idx)}: want {want[f][tuple(idx)]} got {got[f][tuple(idx)]}")
    for f in MW_FIELDS + F64_FIELDS:
        a, b = want[f], got[f]
        if f in MW_FIELDS and mw_exact:
            bad = np.argwhere(a.view(np.uint64) != b.view(np.uint64))
        else:
            bad = np.argwhere(~(np.abs(a - b) <= ABS_TOL + REL_TOL * np.abs(a)))
        for idx in bad[:5]:
            errs.append(f"{label}{f}{tuple(idx)}: want {a[tuple(idx)]!r} got {b[tuple(idx)]!r}")
    return errs


def load_golden():
    import json
    g = json.load(open(os.path.join(HERE, "golden", "format_golden.json")))
    assert g["dtype_itemsize"] == abi.FORMAT_DTYPE.itemsize
    sups = [c["support"] for c in g["random"]]
    want = np.frombuffer(bytes.fromhex("".join(c["record"] for c in g["random"])), dtype=abi.FORMAT_DTYPE).copy()
    return sups, want, g["scipy"]


def scipy_supports(rows, as_bytes):
    """The reference's scipy-derived Mann-Whitney rows as supports: REF values on allele 0, ALT values on
    allele 1, carried by the MAPQ column (as_bytes) or — scaled into [0, 0.5], ranks unchanged — by the
    folded read position."""
    sups = []
    for row in rows:
        vals = np.asarray(row["ref"] + row["alt"], dtype=np.float64)
        n = len(vals)
        sup = simple_support([(0 if i < len(row["ref"]) else 1, 1000 + 7 * i, 0, 1, 10 + i) for i in range(n)])
        if as_bytes:
            assert np.all(vals * 2 == np.round(vals * 2)) and vals.max(initial=0) * 2 < 256
            sup["map_qual"] = (vals * 2).astype(np.uint8)  # one row holds half-integers; doubling keeps the ranks
        else:
            sup["folded_pos"] = vals / 1024.0
        sups.append(sup)
    return sups


def reference_kat_cases():
    """The reference's own known-answer tests for the metrics (tests/caller/variant_support_metrics_test.cpp:
    32-232), as (name, support, field, expectation) with expectation None (= nullopt), a float
    (WithinAbs 1e-6) or a (lo, hi) open interval."""
    E = lambda allele, start, nm, hap, h: (allele, start, nm, hap, h)  # noqa: E731  MakeEvidence's argument order
    return [
        ("FSSE <3 ALT reads", simple_support([E(1, 1000, 0, 1, 100), E(1, 1003, 0, 1, 101)]), "fsse", None),
        ("FSSE one 3bp bin", simple_support([E(1, 999, 0, 1, 100), E(1, 1000, 0, 1, 101), E(1, 1001, 0, 1, 102),
                                            E(1, 999, 0, 1, 103), E(1, 1000, 0, 1, 104)]), "fsse", 0.0),
        ("FSSE diverse starts", simple_support([E(1, i * 100, 0, 1, 200 + i) for i in range(10)]), "fsse", 1.0),
        ("FSSE fraying", simple_support([E(1, 1000 + (i % 3), 0, 1, 100 + i) for i in range(6)]), "fsse", (0.3, 0.4)),
        ("FSSE ignores REF", simple_support([E(0, i * 100, 0, 0, 300 + i) for i in range(10)] +
                                           [E(1, 500, 0, 1, 400), E(1, 600, 0, 1, 401)]), "fsse", None),
        ("AHDD empty REF", simple_support([E(1, 1000, 2, 1, 100), E(1, 1003, 3, 1, 101)]), "ahdd", None),
        ("AHDD empty ALT", simple_support([E(0, 1000, 1, 0, 100), E(0, 1003, 0, 0, 101)]), "ahdd", None),
        ("AHDD equal means", simple_support([E(0, 1000, 2, 0, 100), E(0, 1003, 2, 0, 101), E(1, 1006, 2, 1, 102),
                                            E(1, 1009, 2, 1, 103)]), "ahdd", 0.0),
        ("AHDD ALT worse", simple_support([E(0, 1000, 1, 0, 100), E(0, 1003, 1, 0, 101), E(1, 1006, 5, 1, 102),
                                          E(1, 1009, 5, 1, 103)]), "ahdd", 4.0),
        ("HSE single haplotype", simple_support([E(1, 1000 + i, 0, 1, 100 + i) for i in range(5)], total_haps=1),
         "hse", None),
        ("HSE <3 ALT reads", simple_support([E(1, 1000, 0, 1, 100), E(1, 1003, 0, 2, 101)], total_haps=3), "hse", None),
        ("HSE one path", simple_support([E(1, 1000 + i, 0, 1, 100 + i) for i in range(5)], total_haps=3), "hse", 0.0),
        ("HSE uniform split", simple_support([E(1, 1000, 0, 1, 100), E(1, 1003, 0, 2, 101), E(1, 1006, 0, 3, 102)],
                                            total_haps=3), "hse", 1.0),
        ("HSE ignores REF", simple_support([E(0, 1000, 0, 0, 100), E(0, 1003, 0, 1, 101), E(0, 1006, 0, 2, 102),
                                           E(1, 1009, 0, 1, 103), E(1, 1012, 0, 2, 104)], total_haps=3), "hse", None),
    ]


def check_kats(records):
    errs = []
    for (name, _sup, field, exp), rec in zip(reference_kat_cases(), records):
        has = bool(rec["valid"] & abi.LGR_FMT_HAS[field])
        if exp is None:
            if has:
                errs.append(f"{name}: expected nullopt, got {rec[field]}")
        elif not has:
            errs.append(f"{name}: expected a value, got nullopt")
        elif isinstance(exp, tuple):
            if not exp[0] < rec[field] < exp[1]:
                errs.append(f"{name}: {rec[field]} not in {exp}")
        elif abs(rec[field] - exp) > 1e-6:
            errs.append(f"{name}: want {exp} got {rec[field]}")
    return errs


def check_scipy(rows, records, field):
    errs = []
    for row, rec in zip(rows, records):
        has = bool(rec["valid"] & abi.LGR_FMT_HAS[field])
        if row["expected"] is None:
            if has:
                errs.append(f"{field}: expected nullopt for an empty group")
        elif not has or abs(rec[field] - row["expected"]) > 1e-9:  # the reference's EFFECT_SIZE_TOLERANCE
            errs.append(f"{field}: want {row['expected']} got {rec[field]} (valid={has})")
    return errs


def supports_from_assignments(batch, assign, seed=0, n_samples=2):
    """Genotyper::AddToTable (genotyper.cpp:423-456) on the path's own output: the lgr_assign records of a
    batch become one evidence stream per (variant, sample), in read order.  Read metadata the
    realignment does not carry (insert size, start, MAPQ, SAM flags, sample) is drawn from `seed`;
    mates (reads 2i, 2i+1 of a group) share a name hash, so the dedup has work."""
    rng = np.random.default_rng(seed)
    sups = []
    for g in range(batch.n_groups):
        rb, re = int(batch.grp_read_begin[g]), int(batch.grp_read_begin[g + 1])
        vb, ve = int(batch.grp_var_begin[g]), int(batch.grp_var_begin[g + 1])
        P = int(batch.grp_hap_begin[g + 1] - batch.grp_hap_begin[g])
        R = re - rb
        meta = dict(isz=rng.integers(-600, 600, R) * (rng.random(R) < 0.9), start=rng.integers(1000, 3000, R),
                    mapq=rng.choice([0, 20, 60, 60, 60], R), flags=rng.integers(0, 8, R),
                    sample=np.sort(rng.integers(0, n_samples, R)))
        for v in range(vb, ve):
            lo, hi = int(batch.var_hap_off[v]), int(batch.var_hap_off[v + 1])
            k = max(2, int(batch.var_allele[lo:hi].max()) + 1)
            for s in range(n_samples):
                rows = []
                for i in range(R):
                    a = assign[int(batch.asg_off[rb + i]) + (v - vb)]
                    if meta["sample"][i] != s or not a["assigned"]:
                        continue
                    rows.append((meta["isz"][i], meta["start"][i],
                                 float(a["global_score"]) + float(a["local_score"]) * float(a["local_identity"]),
                                 a["folded_read_pos"], batch.read_name_hash[rb + i - (i % 2)], a["ref_nm"], a["own_hap_nm"],
                                 a["hap_id"], a["allele"], meta["flags"][i], a["base_qual"], meta["mapq"][i]))
                cols = list(zip(*rows)) if rows else [[] for _ in abi.EVIDENCE_FIELDS]
                sup = {name: np.asarray(col, dtype=dt) for (name, dt), col in zip(abi.EVIDENCE_FIELDS, cols)}
                sup.update(n_alleles=k, variant_len=int(batch.var_len[lo]), total_haps=P)
                sups.append(sup)
    return sups


def edge_supports(seed=12):
    """extreme qualities, 2^40 insert sizes / starts, bins around zero, big tie groups, zero variance,
    one read name, single-allele and 8-allele supports, u32 / i32 maxima"""
    rng = np.random.default_rng(seed)

    def sup(n, k, **over):
        s = random_support(rng, n=n, n_alleles=k)
        s.update(over)
        return s

    n = 40
    return [
        sup(1, 2), sup(1, 1), sup(2, 2, allele=np.array([0, 1])), sup(3, 2, allele=np.array([1, 1, 1])),
        sup(n, 2, base_qual=np.full(n, 255)), sup(n, 2, base_qual=np.zeros(n, int)), sup(n, 2, base_qual=np.full(n, 93)),
        sup(n, 2, insert_size=rng.integers(-(1 << 40), 1 << 40, n)), sup(n, 2, aln_start=rng.integers(-(1 << 40), 1 << 40, n)),
        sup(n, 2, aln_start=np.arange(n) - 20),                       # bins around zero: C truncation of start / 3
        sup(n, 2, folded_pos=np.where(np.arange(n) % 2, 0.0, 0.5)),    # two big tie groups
        sup(n, 2, folded_pos=np.full(n, 0.25), map_qual=np.full(n, 60), base_qual=np.full(n, 37)),  # var_u = 0 → 0.0
        sup(n, 2, rname_hash=np.full(n, 7)),                           # one read name: one record per allele survives
        sup(n, 2, allele=np.zeros(n, int)), sup(n, 2, allele=np.ones(n, int)), sup(n, 1, allele=np.zeros(n, int)),
        sup(n, 8, allele=np.arange(n) % 8), sup(n, 3, allele=np.arange(n) % 3, total_haps=1),
        sup(n, 2, aln_score=np.full(n, -1e300)), sup(n, 2, hap_id=np.full(n, 4294967295)),
        sup(n, 2, ref_nm=np.full(n, 4294967295), own_hap_nm=np.full(n, 4294967295), variant_len=2147483647),
    ]


def load_golden_edge():
    import json
    g = json.load(open(os.path.join(HERE, "golden", "format_golden.json")))
    sups = [c["support"] for c in g["edge"]]
    want = np.frombuffer(bytes.fromhex("".join(c["record"] for c in g["edge"])), dtype=abi.FORMAT_DTYPE).copy()
    return sups, want


# ---------------------------------------------------------------------------------------------
# The text a VCF reader sees.  north_star: "and therefore the final VCF records" must be equal.
# VariantCall narrows VariantSupport's f64 metrics to f32 and SampleFormatData prints them with fixed
# precision (reference: src/lancet/caller/variant_call.cpp:160-209 (narrowing, NPBQ = raw PBQ / depth,
# :366-379), sample_format_data.cpp:45-92 (formats: RMQ/NPBQ "%.1F" on f32, SB "{:.3f}", SCA "{:.4f}",
# FLD "{:.1f}", RPCD/BQCD/MQCD/FSSE/HSE "{:.4f}", ASMD/AHDD "{:.3f}", CMLOD "%.4F" on f64 from index 1,
# PL joined by ',', missing optional = ".")).  Those two files need fmt and abseil, which are not in
# this container, so the formatting is restated here: printf-style fixed formatting of the exact binary
# value, which is what both fmt and absl::StrFormat produce.  Fields that VariantCall derives from
# other inputs (GT, DP, SDFC, PRAD, PANG, PDCV) are not part of lgr_format and are left out.
# ---------------------------------------------------------------------------------------------
def render_vcf_fields(rec) -> str:
    k = int(rec["n_alleles"])
    f32 = np.float32

    def fx(v, nd):  # fixed formatting of an f32 value
        return f"%.{nd}f" % float(f32(v))

    def opt(name, nd):
        return fx(rec[name], nd) if int(rec["valid"]) & abi.LGR_FMT_HAS[name] else "."

    ad = [int(rec["fwd"][a]) + int(rec["rev"][a]) for a in range(k)]
    npbq = [float(rec["raw_pbq"][a]) / ad[a] if ad[a] > 0 else 0.0 for a in range(k)]
    parts = [
        ",".join(str(x) for x in ad), ",".join(str(int(rec["fwd"][a])) for a in range(k)),
        ",".join(str(int(rec["rev"][a])) for a in range(k)),
        ",".join(fx(rec["rms_mq"][a], 1) for a in range(k)), ",".join(fx(v, 1) for v in npbq),
        fx(rec["sb"], 3), fx(rec["sca"], 4), opt("fld", 1), opt("rpcd", 4), opt("bqcd", 4), opt("mqcd", 4), opt("asmd", 3),
        ",".join("%.4f" % float(rec["cmlod"][a]) for a in range(1, k)) if k >= 2 else ".",
        opt("fsse", 4), opt("ahdd", 3), opt("hse", 4),
        ",".join(str(int(x)) for x in rec["pl"][:k * (k + 1) // 2]), str(int(rec["gq"])),
    ]
    return ":".join(parts)


def vcf_string_mismatches(want, got, limit=10):
    bad = []
    for i, (w, g) in enumerate(zip(want, got)):
        a, b = render_vcf_fields(w), render_vcf_fields(g)
        if a != b:
            bad.append(f"support {i}: reference {a}\n{' ' * (len(str(i)) + 10)}     ours {b}")
            if len(bad) >= limit:
                break
    return bad
