"""C++ adapter half of the FORMAT-math row: lancet_gpu::EvidenceColumns turns
the realignment's lgr_assign records into the SoA evidence the device consumes (AddToTable,
genotyper.cpp:423-456), lancet_gpu::GpuFormatMetrics runs lgr_format_metrics on it.
 * CPU: the columns against an independent Python model of AddToTable, and the emulated device
   arithmetic on them against the reference's VariantSupport (oracle/_ref) fed record by record;
 * GPU: GpuFormatMetrics::Compute on the same columns against the host build of the same core."""
import ctypes as C

import numpy as np
import pytest

import format_lib as F
import oracle_lib as O
from lancet2_b200 import abi, synth
from test_adapter import lib  # noqa: F401  (module fixture: loads the library, building it when missing)

SNAMES = ("sampleA", "sampleB", "sampleC")


def make_case(seed=23):
    rng = np.random.default_rng(seed)
    groups = synth.make_region_groups(9, ref_len=40_000)[:4] + synth.make_groups(3, 2, n_reads=64, n_haps=4, hap_len=600)
    batch = abi.Batch(groups)
    nr = batch.n_reads
    names = []
    for g in groups:  # mates: reads 2i and 2i+1 of a group share a name → the dedup has work
        names += [g.names[i - (i % 2)] for i in range(len(g.names))]
    meta = dict(sample_id=rng.integers(0, len(SNAMES), nr).astype(np.int32), start0=rng.integers(10_000, 20_000, nr).astype(np.int64),
                isize=(rng.integers(-500, 500, nr) * (rng.random(nr) < 0.9)).astype(np.int64),
                flag=(rng.integers(0, 2, nr) * 0x10 + rng.integers(0, 2, nr) * 0x2).astype(np.uint16),
                mapq=rng.integers(0, 61, nr).astype(np.uint8), softclip=rng.integers(0, 2, nr).astype(np.uint8))
    want, _ = O.oracle_genotype(batch, O.default_params(), n_threads=4)
    return groups, batch, names, meta, want


def call_args(batch, names, meta, want):
    bi = batch.c_struct()
    return [C.byref(bi), b"\0".join(x.encode() for x in names) + b"\0", b"\0".join(s.encode() for s in SNAMES) + b"\0",
            meta["sample_id"].ctypes.data, meta["start0"].ctypes.data, meta["isize"].ctypes.data, meta["flag"].ctypes.data,
            meta["mapq"].ctypes.data, meta["softclip"].ctypes.data, want.assign.ctypes.data], bi


def python_model(groups, batch, names, meta, want):
    """AddToTable, independently: supports in (group, variant, sample first seen) order, records in read order."""
    sups, keys = [], []
    for g_i, g in enumerate(groups):
        r0, r1 = int(batch.grp_read_begin[g_i]), int(batch.grp_read_begin[g_i + 1])
        vb = int(batch.grp_var_begin[g_i])
        P = len(g.haps)
        for v in range(len(g.variants)):
            order, rows = [], {}
            for r in range(r0, r1):
                a = want.assign[batch.asg_off[r] + v]
                if not a["assigned"]:
                    continue
                s = int(meta["sample_id"][r])
                if s not in rows:
                    order.append(s)
                    rows[s] = []
                fl = (abi.LGR_EV_REV if meta["flag"][r] & 0x10 else 0) | (abi.LGR_EV_SOFTCLIP if meta["softclip"][r] else 0) | \
                     (abi.LGR_EV_PROPER_PAIR if meta["flag"][r] & 0x2 else 0)
                rows[s].append((meta["isize"][r], meta["start0"][r],
                                float(a["global_score"]) + float(a["local_score"]) * float(a["local_identity"]),
                                a["folded_read_pos"], abi.x31_hash(names[r]), a["ref_nm"], a["own_hap_nm"], a["hap_id"],
                                a["allele"], fl, a["base_qual"], meta["mapq"][r]))
            lo = int(batch.var_hap_off[vb + v])
            al = batch.var_allele[lo:lo + P]
            k = 1 + max(0, int(al.max()))
            ref_len = int(batch.var_len[lo])
            vlen = max([abs(int(batch.var_len[lo + h]) - ref_len) for h in range(1, P) if al[h] > 0] or [0])
            for s in order:
                sup = {name: np.asarray(col, dtype=dt) for (name, dt), col in zip(abi.EVIDENCE_FIELDS, zip(*rows[s]))}
                sup.update(n_alleles=k, variant_len=vlen, total_haps=P)
                sups.append(sup)
                keys.append((g_i, v, s))
    return sups, keys


@pytest.mark.parametrize("seed", [23, 24, 25])
def test_evidence_columns_match_add_to_table_model(lib, seed):  # noqa: F811
    groups, batch, names, meta, want = make_case(seed)
    lib.lgr_adapter_evidence_columns.argtypes = [C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + [C.c_void_p] * 7 + \
        [C.c_void_p, C.c_int]
    lib.lgr_adapter_evidence_columns.restype = C.POINTER(abi.LgrEvidenceIn)
    sups, keys = python_model(groups, batch, names, meta, want)
    key_out = np.zeros(3 * (len(keys) + 8), dtype=np.int32)
    args, _bi = call_args(batch, names, meta, want)
    p = lib.lgr_adapter_evidence_columns(*args, key_out.ctypes.data, len(keys) + 8)
    assert bool(p)
    ev = p.contents
    model = abi.EvidenceBatch(sups)
    assert (ev.n_supports, ev.n_evidence) == (model.n_supports, model.n_evidence) and ev.n_supports > 10
    assert [tuple(key_out[3 * s:3 * s + 3]) for s in range(len(keys))] == keys

    def col(ptr, n, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()

    for name in ("sup_begin", "sup_n_alleles", "sup_variant_len", "sup_total_haps"):
        want_arr = getattr(model, name)
        assert np.array_equal(col(getattr(ev, name), len(want_arr), want_arr.dtype), want_arr), name
    for name, dt in abi.EVIDENCE_FIELDS:
        assert np.array_equal(col(getattr(ev, name), model.n_evidence, dt), model.cols[name]), name
    # the emulated device arithmetic on the adapter's own columns vs the reference's VariantSupport, record by record
    out = np.zeros(ev.n_supports, dtype=abi.FORMAT_DTYPE)
    assert F.load_emu().emu_format_metrics_ex(p, out.ctypes.data, 1) == 0
    assert int(out["n_kept"].sum()) < model.n_evidence      # mates were deduplicated
    if F.have_ref():
        errs = F.compare_format(F.ref_format(sups), out)
        assert not errs, "\n".join(errs[:20])


@pytest.mark.gpu
def test_gpu_format_metrics_class(lib):  # noqa: F811
    groups, batch, names, meta, want = make_case()
    sups, keys = python_model(groups, batch, names, meta, want)
    lib.lgr_adapter_format_metrics.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + [C.c_void_p] * 7 + \
        [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    lib.lgr_adapter_format_metrics.restype = C.c_int
    out = np.zeros(len(keys), dtype=abi.FORMAT_DTYPE)
    err = C.create_string_buffer(512)
    args, _bi = call_args(batch, names, meta, want)
    n = lib.lgr_adapter_format_metrics(0, *args, out.ctypes.data, len(out), err, len(err))
    assert n == len(keys), err.value.decode()
    rc, emu = F.emu_format(sups)
    assert rc == 0
    errs = F.compare_format(emu, out)
    assert not errs, "\n".join(errs[:20])


@pytest.mark.gpu
def test_gpu_golden_edge_supports():
    from lancet2_b200.format_metrics import GpuFormatMetrics
    sups, want = F.load_golden_edge()
    fmt = GpuFormatMetrics(0)
    try:
        got, _ = fmt.compute(sups)
    finally:
        fmt.close()
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])


@pytest.mark.gpu
def test_gpu_more_supports_than_resident_ctas():
    # beyond 32 CTAs per SM and task (4736 supports on a B200) k_fmt_metrics walks the supports in a grid-stride loop
    from lancet2_b200.format_metrics import GpuFormatMetrics
    rng = np.random.default_rng(321)
    sups = [F.random_support(rng, n=int(rng.integers(0, 24))) for _ in range(6000)]
    rc, want = F.emu_format(sups)
    assert rc == 0
    fmt = GpuFormatMetrics(0)
    try:
        got, _ = fmt.compute(sups)
    finally:
        fmt.close()
    errs = F.compare_format(want, got)
    assert not errs, "\n".join(errs[:20])
