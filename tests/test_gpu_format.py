"""GPU tests of the FORMAT-math row, through the C-ABI (lgr_format_metrics):
k_fmt_dedup / k_fmt_metrics against the golden vectors generated from the reference's own
VariantSupport, its known-answer tests and scipy Mann-Whitney fixture, and against the same
arithmetic compiled for the host.  Tolerances (tests/format_lib.py): counts, PL, GQ and the three
Mann-Whitney effect sizes are exact; the other f64 metrics 1e-9 relative (CUDA's libm differs
from glibc in the last bits, and the sums are warp trees)."""
import numpy as np
import pytest

import format_lib as F
from lancet2_b200 import abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fmt():
    from lancet2_b200.format_metrics import GpuFormatMetrics
    f = GpuFormatMetrics(0)
    yield f
    f.close()


def test_gpu_golden_random_supports(fmt):
    sups, want, _ = F.load_golden()
    got, ms = fmt.compute(sups)
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])
    assert ms > 0.0


def test_gpu_reference_known_answer_cases(fmt):
    got, _ = fmt.compute([c[1] for c in F.reference_kat_cases()])
    errs = F.check_kats(got)
    assert not errs, "\n".join(errs)


@pytest.mark.parametrize("as_bytes,field", [(True, "mqcd"), (False, "rpcd")])
def test_gpu_scipy_mann_whitney_rows(fmt, as_bytes, field):
    _, _, rows = F.load_golden()
    got, _ = fmt.compute(F.scipy_supports(rows, as_bytes))
    errs = F.check_scipy(rows, got, field)
    assert not errs, "\n".join(errs)


def test_gpu_matches_host_build_of_the_same_core(fmt):
    # 3000 supports incl. a few of ~1000 records and 8-allele ones; several grid-stride rounds per warp
    rng = np.random.default_rng(123)
    sups = [F.random_support(rng) for _ in range(2960)]
    sups += [F.random_support(rng, n=int(rng.integers(600, 1200)), n_alleles=2) for _ in range(30)]
    sups += [F.random_support(rng, n=200, n_alleles=8) for _ in range(10)]
    rc, want = F.emu_format(sups)
    assert rc == 0
    got, _ = fmt.compute(sups)
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])
    got2, _ = fmt.compute(sups)          # run-to-run determinism (fixed reduction trees)
    assert got.tobytes() == got2.tobytes()
    if F.have_ref():
        errs = F.compare_format(F.ref_format(sups[-60:]), got[-60:])
        assert not errs, "\n".join(errs[:20])


def test_gpu_realign_then_format(fmt):
    # the hot path feeding its consumer: GPU assignments → AddToTable (host) → GPU FORMAT math
    from lancet2_b200.realign import GpuRealigner
    gpu = GpuRealigner(0)
    try:
        batch = abi.Batch(synth.make_groups(42, 6, read_len=150, hap_len=800, n_haps=4, n_reads=120))
        res, _ = gpu.genotype_batch(batch)
    finally:
        gpu.close()
    sups = F.supports_from_assignments(batch, res.assign, seed=3)
    rc, want = F.emu_format(sups)
    assert rc == 0
    got, _ = fmt.compute(sups)
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])
    if F.have_ref():
        errs = F.compare_format(F.ref_format(sups), got)
        assert not errs, "\n".join(errs[:20])


def test_gpu_format_argument_errors(fmt):
    rng = np.random.default_rng(6)
    got, _ = fmt.compute([])
    assert len(got) == 0
    got, _ = fmt.compute([F.random_support(rng, n=0, n_alleles=2)])
    assert got[0]["n_kept"] == 0 and got[0]["valid"] == 0
    bad = F.random_support(rng, n=5, n_alleles=2)
    bad["allele"] = np.array([0, 1, 2, 0, 1])
    with pytest.raises(RuntimeError, match="allele index"):
        fmt.compute([bad])
    # a site with more alleles than a record holds (STR loci): flagged per support, the batch is not failed
    wide = F.random_support(rng, n=40, n_alleles=11)
    ok = [F.random_support(rng, n=30, n_alleles=3), F.random_support(rng, n=25, n_alleles=8)]
    got, _ = fmt.compute([ok[0], wide, ok[1]])
    assert fmt.last_rc == abi.LGR_E_PARTIAL
    assert got[1]["valid"] == abi.LGR_FMT_WIDE and got[1]["n_alleles"] == 11 and got[1]["n_kept"] == 0
    alone, _ = fmt.compute(ok)
    assert fmt.last_rc == 0 and got[0].tobytes() == alone[0].tobytes() and got[2].tobytes() == alone[1].tobytes()


def test_gpu_vcf_text_equals_the_references(fmt):
    """the device records rendered the way VariantCall / SampleFormatData print them must be the same
    text as the reference's: golden supports always, 10^4 random ones when oracle/_ref travelled"""
    sups, want, _ = F.load_golden()
    got, _ = fmt.compute(sups)
    bad = F.vcf_string_mismatches(want, got)
    assert not bad, "\n".join(bad)
    if F.have_ref():
        rng = np.random.default_rng(2026)
        rs = [F.random_support(rng) for _ in range(10_000)]
        got, _ = fmt.compute(rs)
        bad = F.vcf_string_mismatches(F.ref_format(rs), got)
        assert not bad, "\n".join(bad)
