"""CPU tests of the FORMAT-math row: the arithmetic the device runs
(lancet2_b200/csrc/lgr_format.cuh, compiled by g++ with the warp emulated) against the
reference's own VariantSupport — golden vectors generated from oracle/_ref, the reference's
known-answer tests, its scipy-derived Mann-Whitney fixture, and oracle/_ref live when present."""
import ctypes as C

import numpy as np
import pytest

import format_lib as F
from lancet2_b200 import abi


def test_struct_layouts_match_the_header():
    # sizes printed by a C program over include/lancet_gpu_realign.h: 592 / 144
    assert abi.FORMAT_DTYPE.itemsize == 592
    assert C.sizeof(abi.LgrEvidenceIn) == 144
    assert abi.FORMAT_DTYPE.fields["fwd"][1] == 32 * 8 + 10 * 8
    assert abi.FORMAT_DTYPE.fields["gq"][1] == 32 * 8 + 10 * 8 + 3 * 32 + 36 * 4


def test_golden_random_supports():
    sups, want, _ = F.load_golden()
    rc, got = F.emu_format(sups)
    assert rc == 0
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])


def test_reference_known_answer_cases():
    cases = F.reference_kat_cases()
    rc, got = F.emu_format([c[1] for c in cases])
    assert rc == 0
    errs = F.check_kats(got)
    assert not errs, "\n".join(errs)


@pytest.mark.parametrize("as_bytes,field", [(True, "mqcd"), (False, "rpcd")])
def test_scipy_mann_whitney_rows(as_bytes, field):
    _, _, rows = F.load_golden()
    rc, got = F.emu_format(F.scipy_supports(rows, as_bytes))
    assert rc == 0
    errs = F.check_scipy(rows, got, field)
    assert not errs, "\n".join(errs)


@pytest.mark.skipif(not F.have_ref(), reason="oracle/_ref not built (needs the reference tree)")
def test_live_reference_random_supports():
    rng = np.random.default_rng(99)
    sups = [F.random_support(rng) for _ in range(600)]
    sups += [F.random_support(rng, n=int(rng.integers(300, 900)), n_alleles=int(rng.integers(2, 4))) for _ in range(8)]
    rc, got = F.emu_format(sups)
    assert rc == 0
    errs = F.compare_format(F.ref_format(sups), got)
    assert not errs, "\n".join(errs[:20])


def test_dedup_is_first_seen_per_allele():
    # same read-name hash: twice on allele 1 (second dropped, even though its strand differs), once on allele 0 (kept)
    sup = F.simple_support([(1, 1000, 0, 1, 7), (1, 1010, 0, 1, 7), (0, 1020, 0, 0, 7), (1, 1030, 0, 1, 8)])
    sup["flags"] = np.array([0, abi.LGR_EV_REV, 0, abi.LGR_EV_REV])
    rc, got = F.emu_format([sup])
    assert rc == 0
    r = got[0]
    assert (r["n_kept"], r["fwd"][0], r["rev"][0], r["fwd"][1], r["rev"][1]) == (3, 1, 0, 1, 1)


def test_results_do_not_depend_on_batch_composition():
    rng = np.random.default_rng(5)
    sups = [F.random_support(rng) for _ in range(40)]
    _, together = F.emu_format(sups)
    for i in (0, 7, 39):
        _, alone = F.emu_format([sups[i]])
        assert alone[0].tobytes() == together[i].tobytes()


def test_task_split_equals_the_single_call():
    rng = np.random.default_rng(8)
    sups = [F.random_support(rng) for _ in range(200)]
    _, split = F.emu_format(sups, split_tasks=True)
    _, whole = F.emu_format(sups, split_tasks=False)
    assert split.tobytes() == whole.tobytes()


def test_empty_support_and_limits():
    rng = np.random.default_rng(6)
    rc, got = F.emu_format([F.random_support(rng, n=0, n_alleles=2)])
    assert rc == 0 and got[0]["n_kept"] == 0 and got[0]["valid"] == 0 and list(got[0]["pl"][:3]) == [0, 0, 0]
    bad = F.random_support(rng, n=5, n_alleles=2)
    bad["allele"] = np.array([0, 1, 2, 0, 1])
    assert F.emu_format([bad])[0] == -1      # LGR_E_ARG: allele index >= K
    bad = F.random_support(rng, n=5, n_alleles=2)
    bad["n_alleles"] = 9
    assert F.emu_format([bad])[0] == -4      # LGR_E_LIMIT


@pytest.mark.skipif(not F.have_ref(), reason="oracle/_ref not built (needs the reference tree)")
def test_supports_built_from_the_paths_own_assignments():
    # realignment oracle → AddToTable → FORMAT math: emulated device arithmetic vs the reference's VariantSupport
    import oracle_lib as O
    from lancet2_b200 import synth
    batch = abi.Batch(synth.make_groups(42, 6, read_len=150, hap_len=800, n_haps=4, n_reads=120))
    params = abi.LgrParams()
    abi.load_library().lgr_default_params(C.byref(params))
    res, _ = O.oracle_genotype(batch, params, n_threads=4)
    sups = F.supports_from_assignments(batch, res.assign, seed=3)
    assert sum(len(s["allele"]) for s in sups) > 300
    rc, got = F.emu_format(sups)
    assert rc == 0
    want = F.ref_format(sups)
    assert int(want["n_kept"].sum()) < sum(len(s["allele"]) for s in sups)  # the mate dedup removed something
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])


def test_device_path_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from lancet2_b200.format_metrics import GpuFormatMetrics
    with pytest.raises(RuntimeError, match="no CUDA device|no usable CUDA device"):
        GpuFormatMetrics(0)


@pytest.mark.skipif(not F.have_ref(), reason="oracle/_ref not built (needs the reference tree)")
def test_edge_supports_against_live_reference():
    cases = F.edge_supports()
    rc, got = F.emu_format(cases)
    assert rc == 0
    errs = F.compare_format(F.ref_format(cases), got)
    assert not errs, "\n".join(errs[:20])


def test_golden_edge_supports():
    sups, want = F.load_golden_edge()
    assert len(sups) >= 20
    rc, got = F.emu_format(sups)
    assert rc == 0
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])


def test_sort_build_equals_scan_build():
    """-DLGR_FMT_SORT (the shipped device build since round 2): ranks and bins from a
    shared-memory bitonic sort instead of the O(n^2) scans.  The Mann-Whitney statistics are
    integers, so RPCD must keep its bits; the entropies add their terms in sorted-bin order
    (tolerance).  Supports beyond the 2048-record cap take the scan in both builds."""
    rng = np.random.default_rng(77)
    sups = [F.random_support(rng) for _ in range(300)] + F.edge_supports()
    sups += [F.random_support(rng, n=n, n_alleles=2) for n in (127, 128, 129, 1000, 2047, 2048, 2049, 2500)]
    sups += [dict(F.random_support(rng, n=50, n_alleles=2), folded_pos=np.where(np.arange(50) % 2, -0.0, 0.0))]
    rc, scan = F.emu_format(sups)
    rc2, sort = F.emu_format_sort(sups)
    assert rc == 0 and rc2 == 0
    for f in ("mqcd", "rpcd", "bqcd"):
        assert scan[f].tobytes() == sort[f].tobytes(), f
    errs = F.compare_format(scan, sort)
    assert not errs, "\n".join(errs[:20])
    sg, want, _ = F.load_golden()
    errs = F.compare_format(want, F.emu_format_sort(sg)[1])
    assert not errs, "\n".join(errs[:20])


def test_scan_variant_library_exports_the_same_entry_points():
    import os
    path = os.path.join(os.path.dirname(abi.LIB_PATH), "liblgr_format_scan.so")
    if not os.path.exists(path):
        pytest.skip("variant library not built")
    lib = C.CDLL(path)
    for sym in ("lgr_format_create", "lgr_format_destroy", "lgr_format_last_error", "lgr_format_metrics"):
        getattr(lib, sym)
    assert not hasattr(lib, "lgr_genotype_batch")   # only the FORMAT entry points live there


def test_checker_is_not_vacuous():
    # compare_format must flag a one-unit PL change, a flipped valid bit, a last-bit change of a Cohen's d and a 1e-6 drift of an entropy
    sups, want, _ = F.load_golden()
    for field, mutate in (("pl", lambda r: r["pl"].__setitem__(1, r["pl"][1] + 1)),
                          ("valid", lambda r: r.__setitem__("valid", r["valid"] ^ 32)),
                          ("rpcd", lambda r: r.__setitem__("rpcd", np.nextafter(r["rpcd"], 1.0))),
                          ("fsse", lambda r: r.__setitem__("fsse", r["fsse"] + 1e-6)),
                          ("fwd", lambda r: r["fwd"].__setitem__(0, r["fwd"][0] + 1))):
        got = want.copy()
        idx = next(i for i in range(len(got)) if got[i]["valid"] & abi.LGR_FMT_HAS["fsse"] and got[i]["rpcd"] != 0.0)
        mutate(got[idx])
        errs = F.compare_format(want, got)
        assert errs and any(field in e for e in errs), field


def test_permutation_invariance_without_duplicates():
    """oracle-free property at a size no reference run is needed for: with unique read names the order of
    a support's records only changes the association of the f64 sums — integers and the Mann-Whitney
    statistics keep their bits, the rest stays within the tolerance"""
    rng = np.random.default_rng(4)
    n = 5000                                  # beyond the sort variant's cap, far beyond any golden support
    sup = F.random_support(rng, n=n, n_alleles=3, dup_frac=0.0)
    sup["rname_hash"] = np.arange(n, dtype=np.uint32)
    perm = rng.permutation(n)
    shuffled = {k: (np.asarray(v)[perm] if not np.isscalar(v) and len(np.shape(v)) == 1 else v) for k, v in sup.items()}
    _, a = F.emu_format([sup])
    _, b = F.emu_format([shuffled])
    assert a[0]["n_kept"] == n
    errs = F.compare_format(a, b)
    assert not errs, "\n".join(errs)
    _, c = F.emu_format_sort([shuffled])
    errs = F.compare_format(a, c)
    assert not errs, "\n".join(errs)


@pytest.mark.skipif(not F.have_ref(), reason="oracle/_ref not built (needs the reference tree)")
def test_fuzz_up_to_eight_alleles_against_live_reference():
    rng = np.random.default_rng(2025)
    sups = []
    for _ in range(3000):
        k, n = int(rng.integers(1, 9)), int(rng.integers(0, 200))
        s = F.random_support(rng, n=n, n_alleles=k, dup_frac=float(rng.random() * 0.8))
        if k > 1 and n:
            s["allele"] = rng.integers(0, k, n)
        s["total_haps"], s["variant_len"] = int(rng.integers(0, 70)), int(rng.integers(0, 400))
        sups.append(s)
    want = F.ref_format(sups)
    for got in (F.emu_format(sups)[1], F.emu_format_sort(sups)[1]):   # scan build and sort build
        errs = F.compare_format(want, got)
        assert not errs, "\n".join(errs[:20])


def test_vcf_text_of_golden_supports_equals_the_references():
    """the FORMAT text a VCF reader sees (f32 narrowing + fixed precision, format_lib.render_vcf_fields):
    the golden records made by the reference's own VariantSupport against the host build of the core"""
    sups, want, _ = F.load_golden()
    rc, got = F.emu_format(sups)
    assert rc == 0
    assert not F.vcf_string_mismatches(want, got)
    esups, ewant = F.load_golden_edge()
    rc, egot = F.emu_format(esups)
    assert rc == 0 and not F.vcf_string_mismatches(ewant, egot)


@pytest.mark.skipif(not F.have_ref(), reason="oracle/_ref not built (needs the reference tree)")
def test_vcf_text_of_ten_thousand_random_supports_equals_the_live_reference():
    """1e-9 relative agreement does not by itself exclude a flipped rounding boundary in the printed
    text: 10^4 random supports rendered from the live reference and from the core must be identical
    strings"""
    rng = np.random.default_rng(2026)
    sups = [F.random_support(rng) for _ in range(10_000)]
    want = F.ref_format(sups)
    rc, got = F.emu_format(sups)
    assert rc == 0
    bad = F.vcf_string_mismatches(want, got)
    assert not bad, "\n".join(bad)
