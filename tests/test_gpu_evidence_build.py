"""SURVEY.md §8f #2, second step — AddToTable on the device (k_evidence_count / _scan / _scatter behind
lgr_format_from_assign): the evidence columns built on the GPU from lgr_assign records must be byte-equal to
the host AddToTable (an independent Python model, which tests/test_zz_format_adapter.py holds against the C++
EvidenceColumns::AppendJob), and the FORMAT records computed from them equal to lgr_format_metrics on the
host-built columns — with the records uploaded, and read in place from the realignment context."""
import numpy as np
import pytest

import format_lib as F
from lancet2_b200 import abi, synth
from lancet2_b200.format_metrics import GpuFormatMetrics
from lancet2_b200.realign import GpuRealigner
from test_zz_format_adapter import SNAMES, make_case, python_model

pytestmark = pytest.mark.gpu


def variant_tables(groups, batch):
    """K = 1 + highest ALT index and max |ALT length - REF length| per variant, as the host adapter derives them."""
    k, vlen = [], []
    for g_i, g in enumerate(groups):
        P, vb = len(g.haps), int(batch.grp_var_begin[g_i])
        for v in range(len(g.variants)):
            lo = int(batch.var_hap_off[vb + v])
            al = batch.var_allele[lo:lo + P]
            ref_len = int(batch.var_len[lo])
            k.append(1 + max(0, int(al.max())))
            vlen.append(max([abs(int(batch.var_len[lo + h]) - ref_len) for h in range(1, P) if al[h] > 0] or [0]))
    return np.asarray(k, np.int32), np.asarray(vlen, np.int32)


def local_keys(keys, batch):
    """the library numbers variants across the batch; the AddToTable model numbers them inside their group"""
    return [(int(g), int(v) - int(batch.grp_var_begin[g]), int(s)) for g, v, s in keys]


def run_from_assign(fmt, batch, meta, k, vlen, **kw):
    return fmt.from_assign(batch, n_samples=len(SNAMES), sample_id=meta["sample_id"], start0=meta["start0"], isize=meta["isize"],
                           sam_flag=meta["flag"], mapq=meta["mapq"], softclip=meta["softclip"], var_n_alleles=k, var_len=vlen, **kw)


@pytest.mark.parametrize("seed", [23, 24])
def test_device_columns_equal_host_add_to_table(seed):
    groups, batch, names, meta, want = make_case(seed)
    batch.read_name_hash[:batch.n_reads] = [abi.x31_hash(n) for n in names]  # mates share a name: the dedup has work
    sups, keys = python_model(groups, batch, names, meta, want)
    model = abi.EvidenceBatch(sups)
    k, vlen = variant_tables(groups, batch)
    fmt = GpuFormatMetrics(0)
    try:
        got, got_keys, ms = run_from_assign(fmt, batch, meta, k, vlen, host_assign=want.assign[:batch.n_assign])
        cols = fmt.debug_evidence()
        ref, _ = fmt.compute(model)
    finally:
        fmt.close()
    assert local_keys(got_keys, batch) == keys and len(keys) > 10 and ms > 0
    for name in ("sup_begin", "sup_n_alleles", "sup_variant_len", "sup_total_haps"):
        assert np.array_equal(cols[name], getattr(model, name)), name
    for name, _dt in abi.EVIDENCE_FIELDS:
        assert cols[name].tobytes() == model.cols[name].tobytes(), name
    assert got.tobytes() == ref.tobytes()                       # same columns, same kernels: the same bits
    assert int(got["n_kept"].sum()) < model.n_evidence


def test_resident_assignments_never_visit_the_host():
    groups, batch, names, meta, want = make_case(31)
    sups, keys = python_model(groups, batch, names, meta, want)   # from the ORACLE's assignments
    k, vlen = variant_tables(groups, batch)
    gpu, fmt = GpuRealigner(0), GpuFormatMetrics(0)
    try:
        gpu.upload_packed(abi.PackedBatch(groups, gpu.lib))
        gpu.run_resident()
        dev, n = gpu.resident_assign()
        assert dev != 0 and n == batch.n_assign
        batch.read_name_hash[:batch.n_reads] = [abi.x31_hash(nm) for nm in names]
        got, got_keys, _ = run_from_assign(fmt, batch, meta, k, vlen, dev_assign=dev)
        batch_model = abi.EvidenceBatch(sups)
        ref, _ = fmt.compute(batch_model)
    finally:
        fmt.close()
        gpu.close()
    assert local_keys(got_keys, batch) == keys
    assert got.tobytes() == ref.tobytes()
    if F.have_ref():                                               # and against the reference's own VariantSupport
        errs = F.compare_format(F.ref_format(sups), got)
        assert not errs, "\n".join(errs[:20])


def test_four_samples_empty_groups_and_argument_errors():
    groups = synth.make_tile_groups("cfg4", 5, 0, ref_len=30_000)[:6]
    batch = abi.Batch(groups)
    rng = np.random.default_rng(4)
    nr = batch.n_reads
    meta = dict(sample_id=rng.integers(0, 4, nr).astype(np.int32), start0=rng.integers(0, 1 << 40, nr).astype(np.int64),
                isize=rng.integers(-(1 << 40), 1 << 40, nr).astype(np.int64), flag=rng.integers(0, 1 << 12, nr).astype(np.uint16),
                mapq=rng.integers(0, 256, nr).astype(np.uint8), softclip=rng.integers(0, 2, nr).astype(np.uint8))
    k, vlen = variant_tables(groups, batch)
    gpu, fmt = GpuRealigner(0), GpuFormatMetrics(0)
    try:
        res, _ = gpu.genotype_batch(batch)
        kw = dict(n_samples=4, sample_id=meta["sample_id"], start0=meta["start0"], isize=meta["isize"], sam_flag=meta["flag"],
                  mapq=meta["mapq"], softclip=meta["softclip"], var_n_alleles=k, var_len=vlen)
        got, keys, _ = fmt.from_assign(batch, host_assign=res.assign[:batch.n_assign], **kw)
        cols = fmt.debug_evidence()
        # every assigned record lands in exactly one support, in read order inside it
        assert cols["sup_begin"][-1] == int(res.assign[:batch.n_assign]["assigned"].astype(bool).sum())
        assert len(set(map(tuple, keys.tolist()))) == len(keys)
        with pytest.raises(RuntimeError):
            fmt.from_assign(batch, host_assign=res.assign[:batch.n_assign], **{**kw, "n_samples": 33})
        with pytest.raises(RuntimeError):
            fmt.from_assign(batch, host_assign=res.assign[:batch.n_assign], **{**kw, "n_samples": 2})  # sample ids 2, 3 out of range
    finally:
        fmt.close()
        gpu.close()
