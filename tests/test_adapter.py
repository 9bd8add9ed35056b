"""C++ Genotyper adapter (lancet2_b200/host/gpu_genotyper.{h,cpp}):
 * CPU: its AddEvidence restatement against golden dumps made by the reference's own
   VariantSupport::AddEvidence (tests/golden/evidence_golden.json);
 * GPU: Genotype()/GenotypeMany() end to end — evidence per (variant, sample, allele), in the
   reference's read order, must equal what the oracle's assignments give through AddToTable."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from lancet2_b200 import abi, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "evidence_golden.json")
FIELDS = [("isize", np.int64), ("start", np.int64), ("aln", np.float64), ("fold", np.float64), ("hash", np.uint32),
          ("ref_nm", np.uint32), ("own_nm", np.uint32), ("hap_id", np.uint32), ("allele", np.uint8), ("rev", np.uint8),
          ("bq", np.uint8), ("mapq", np.uint8), ("softclip", np.uint8), ("proper", np.uint8)]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lb = abi.load_library()
    lb.lgr_adapter_add_evidence_dump.argtypes = [C.c_int] + [C.c_void_p] * 14 + [C.c_char_p, C.c_longlong]
    lb.lgr_adapter_add_evidence_dump.restype = C.c_int
    lb.lgr_adapter_genotype_dump.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + \
        [C.c_void_p] * 6 + [C.c_char_p, C.c_longlong]
    lb.lgr_adapter_genotype_dump.restype = C.c_int
    lb.lgr_adapter_batcher_dump.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + \
        [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_longlong]
    lb.lgr_adapter_batcher_dump.restype = C.c_int
    return lb


def add_evidence_dump(lib, st):
    arrs = [np.ascontiguousarray(st[k], dtype=dt) for k, dt in FIELDS]
    buf = C.create_string_buffer(1 << 22)
    n = lib.lgr_adapter_add_evidence_dump(len(arrs[0]), *[a.ctypes.data for a in arrs], buf, len(buf))
    assert n >= 0
    return buf.value.decode()


def test_add_evidence_matches_reference_golden(lib):
    cases = json.load(open(GOLD))["cases"]
    assert len(cases) >= 50
    for c in cases:
        assert add_evidence_dump(lib, c["stream"]) == c["dump"]


def expected_evidence(lib, batch, groups, names, sample_id, start0, isize, flag, mapq, softclip, want,
                      snames=("normal", "tumor")):
    """oracle assignments → AddToTable (genotyper.cpp:423-456) → the reference-pinned AddEvidence dump"""
    expect = []
    for g_i, g in enumerate(groups):
        r0, r1 = batch.grp_read_begin[g_i], batch.grp_read_begin[g_i + 1]
        for v in range(len(g.variants)):
            order, streams = [], {}
            for r in range(r0, r1):
                a = want.assign[batch.asg_off[r] + v]
                if not a["assigned"]:
                    continue
                s = int(sample_id[r])
                if s not in streams:
                    order.append(s)
                    streams[s] = {k: [] for k, _ in FIELDS}
                st = streams[s]
                st["isize"].append(int(isize[r]))
                st["start"].append(int(start0[r]))
                st["aln"].append(float(a["global_score"]) + float(a["local_score"]) * float(a["local_identity"]))
                st["fold"].append(float(a["folded_read_pos"]))
                st["hash"].append(abi.x31_hash(names[r]))
                st["ref_nm"].append(int(a["ref_nm"]))
                st["own_nm"].append(int(a["own_hap_nm"]))
                st["hap_id"].append(int(a["hap_id"]))
                st["allele"].append(int(a["allele"]))
                st["rev"].append(1 if flag[r] & 0x10 else 0)
                st["bq"].append(int(a["base_qual"]))
                st["mapq"].append(int(mapq[r]))
                st["softclip"].append(int(softclip[r]))
                st["proper"].append(1 if flag[r] & 0x2 else 0)
            for s in order:
                for line in add_evidence_dump(lib, streams[s]).splitlines():
                    al, rest = line.split("|", 1)
                    expect.append(f"G{g_i} V{v} S{snames[s]} {al}|{rest}")
    return expect


def test_host_logic_packing_and_add_to_table_without_gpu(lib):
    """CPU-only: PackedJob/PackedBatch must reproduce the batch's SoA arrays exactly, and
    PackedBatch::BuildResult (AddToTable) fed with the ORACLE's assignments must give the evidence
    the reference-pinned AddEvidence gives — the adapter's host half, no device involved."""
    rng = np.random.default_rng(17)
    groups = synth.make_region_groups(9, ref_len=40_000)[:5] + synth.make_groups(3, 2, n_reads=48, n_haps=4, hap_len=600)
    batch = abi.Batch(groups)
    nr = batch.n_reads
    names = [nm for g in groups for nm in g.names]
    # four samples (BASELINE configs[3]: multi-sample calling): sample identity only matters to AddToTable
    snames = ("sampleA", "sampleB", "sampleC", "sampleD")
    sample_id = rng.integers(0, 4, nr).astype(np.int32)
    start0 = rng.integers(10_000, 20_000, nr).astype(np.int64)
    isize = (rng.integers(-500, 500, nr) * (rng.random(nr) < 0.9)).astype(np.int64)
    flag = (rng.integers(0, 2, nr) * 0x10 + rng.integers(0, 2, nr) * 0x2).astype(np.uint16)
    mapq = rng.integers(0, 61, nr).astype(np.uint8)
    softclip = rng.integers(0, 2, nr).astype(np.uint8)
    want, _ = O.oracle_genotype(batch, O.default_params(), n_threads=4)
    lib.lgr_adapter_host_logic_dump.argtypes = [C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + [C.c_void_p] * 7 + \
        [C.POINTER(C.c_longlong), C.c_char_p, C.c_longlong]
    lib.lgr_adapter_host_logic_dump.restype = C.c_int
    bi = batch.c_struct()
    buf = C.create_string_buffer(64 << 20)
    bad = C.c_longlong(-1)
    n = lib.lgr_adapter_host_logic_dump(C.byref(bi), b"\0".join(x.encode() for x in names) + b"\0",
                                        b"\0".join(s.encode() for s in snames) + b"\0",
                                        sample_id.ctypes.data, start0.ctypes.data, isize.ctypes.data, flag.ctypes.data,
                                        mapq.ctypes.data, softclip.ctypes.data, want.assign.ctypes.data, C.byref(bad), buf, len(buf))
    assert n >= 0, buf.value.decode()
    assert bad.value == 0
    got = buf.value.decode().splitlines()
    expect = expected_evidence(lib, batch, groups, names, sample_id, start0, isize, flag, mapq, softclip, want, snames)
    assert len(got) == len(expect) and len(got) > 10
    assert got == expect


@pytest.mark.gpu
def test_adapter_genotype_matches_oracle_evidence(lib):
    rng = np.random.default_rng(5)
    groups = synth.make_region_groups(9, ref_len=40_000)[:6] + synth.make_groups(3, 2, n_reads=64, n_haps=5, hap_len=700)
    batch = abi.Batch(groups)
    nr = batch.n_reads
    names = [nm for g in groups for nm in g.names]
    sample_id = np.asarray([0 if nm.startswith("n") else 1 for nm in names], dtype=np.int32)
    start0 = rng.integers(10_000, 20_000, nr).astype(np.int64)
    isize = (rng.integers(-500, 500, nr) * (rng.random(nr) < 0.9)).astype(np.int64)
    flag = (rng.integers(0, 2, nr) * 0x10 + rng.integers(0, 2, nr) * 0x2).astype(np.uint16)
    mapq = rng.integers(0, 61, nr).astype(np.uint8)
    softclip = rng.integers(0, 2, nr).astype(np.uint8)
    bi = batch.c_struct()
    buf = C.create_string_buffer(64 << 20)
    n = lib.lgr_adapter_genotype_dump(0, C.byref(bi), b"\0".join(x.encode() for x in names) + b"\0", b"normal\0tumor\0",
                                      sample_id.ctypes.data, start0.ctypes.data, isize.ctypes.data, flag.ctypes.data,
                                      mapq.ctypes.data, softclip.ctypes.data, buf, len(buf))
    assert n >= 0, buf.value.decode()
    got = buf.value.decode().splitlines()

    prm = O.default_params()
    want, _ = O.oracle_genotype(batch, prm)
    expect = expected_evidence(lib, batch, groups, names, sample_id, start0, isize, flag, mapq, softclip, want)
    assert len(got) == len(expect) and len(got) > 10
    assert got == expect


@pytest.mark.gpu
def test_cross_thread_batcher_equals_synchronous_adapter(lib):
    """GenotypeBatcher: every group issued as its own blocking Genotype() call from
    8 worker threads must give exactly the evidence the one-shot GenotypeMany gives, and the
    calls must actually have shared device batches."""
    rng = np.random.default_rng(11)
    groups = synth.make_region_groups(5, ref_len=60_000)[:24] + synth.make_groups(8, 8, n_reads=96, n_haps=4, hap_len=800)
    batch = abi.Batch(groups)
    nr = batch.n_reads
    names = [nm for g in groups for nm in g.names]
    sample_id = np.asarray([0 if nm.startswith("n") else 1 for nm in names], dtype=np.int32)
    start0 = rng.integers(10_000, 20_000, nr).astype(np.int64)
    isize = (rng.integers(-500, 500, nr) * (rng.random(nr) < 0.9)).astype(np.int64)
    flag = (rng.integers(0, 2, nr) * 0x10 + rng.integers(0, 2, nr) * 0x2).astype(np.uint16)
    mapq = rng.integers(0, 61, nr).astype(np.uint8)
    softclip = rng.integers(0, 2, nr).astype(np.uint8)
    bi = batch.c_struct()
    nm_blob = b"\0".join(x.encode() for x in names) + b"\0"
    args = (sample_id.ctypes.data, start0.ctypes.data, isize.ctypes.data, flag.ctypes.data, mapq.ctypes.data, softclip.ctypes.data)
    buf1 = C.create_string_buffer(64 << 20)
    n1 = lib.lgr_adapter_genotype_dump(0, C.byref(bi), nm_blob, b"normal\0tumor\0", *args, buf1, len(buf1))
    assert n1 > 0, buf1.value.decode()
    buf2 = C.create_string_buffer(64 << 20)
    counters = np.zeros(17, dtype=np.uint64)
    rounds = 3
    # (i) every blocking caller through the batcher thread (LGR_BATCHER_DIRECT=0): the calls must be coalesced
    os.environ["LGR_BATCHER_DIRECT"] = "0"
    try:
        n2 = lib.lgr_adapter_batcher_dump(0, C.byref(bi), nm_blob, b"normal\0tumor\0", *args, 8, rounds, 1, 1, counters.ctypes.data,
                                          buf2, len(buf2))
    finally:
        del os.environ["LGR_BATCHER_DIRECT"]
    assert n2 > 0, buf2.value.decode()
    assert buf2.value == buf1.value
    batches, jobs, pairs, max_jobs = (int(x) for x in counters[:4])
    assert jobs == rounds * len(groups) and pairs == rounds * batch.n_pairs
    assert batches < jobs and max_jobs > 1, (batches, jobs, max_jobs)  # calls were coalesced
    # (ii) the default: up to eight blocking callers run on a device context of their own, the rest are coalesced —
    # with 12 threads both paths are in use, and the evidence is the same
    buf2b = C.create_string_buffer(64 << 20)
    n2b = lib.lgr_adapter_batcher_dump(0, C.byref(bi), nm_blob, b"normal\0tumor\0", *args, 12, rounds, 1, 1, counters.ctypes.data,
                                       buf2b, len(buf2b))
    assert n2b > 0, buf2b.value.decode()
    assert buf2b.value == buf1.value
    assert int(counters[1]) == rounds * len(groups) and int(counters[2]) == rounds * batch.n_pairs
    # Enqueue/Collect: two workers with 16 groups in flight each — same evidence, fewer device batches
    buf3 = C.create_string_buffer(64 << 20)
    n3 = lib.lgr_adapter_batcher_dump(0, C.byref(bi), nm_blob, b"normal\0tumor\0", *args, 2, rounds, 16, 1, counters.ctypes.data,
                                      buf3, len(buf3))
    assert n3 > 0, buf3.value.decode()
    assert buf3.value == buf1.value
    assert int(counters[0]) < batches and int(counters[3]) >= 8, counters
    # GenotypeDispatcher over every GPU of the box (one on the default test box): same evidence, and with
    # more than one device every device gets a share of the payloads
    import torch
    n_dev = min(torch.cuda.device_count(), 8)
    buf4 = C.create_string_buffer(64 << 20)
    n4 = lib.lgr_adapter_batcher_dump(0, C.byref(bi), nm_blob, b"normal\0tumor\0", *args, 4, rounds, 8, n_dev, counters.ctypes.data,
                                      buf4, len(buf4))
    assert n4 > 0, buf4.value.decode()
    assert buf4.value == buf1.value
    per_dev = [int(x) for x in counters[9:9 + n_dev]]
    assert sum(per_dev) == rounds * len(groups) and all(x > 0 for x in per_dev), per_dev


def _meta(rng, groups, batch):
    """per-read metadata as a function of the read NAME, so that a read keeps its metadata whichever
    other groups share the batch"""
    names = [nm for g in groups for nm in g.names]
    hs = np.asarray([abi.x31_hash(nm) for nm in names], dtype=np.uint64)
    sample_id = np.asarray([0 if nm.startswith("n") else 1 for nm in names], dtype=np.int32)
    start0 = (10_000 + hs % 10_000).astype(np.int64)
    isize = ((hs // 7 % 1000).astype(np.int64) - 500) * (hs % 10 != 0)
    flag = ((hs // 3 % 2) * 0x10 + (hs // 5 % 2) * 0x2).astype(np.uint16)
    mapq = (hs // 11 % 61).astype(np.uint8)
    softclip = (hs // 13 % 2).astype(np.uint8)
    blob = b"\0".join(x.encode() for x in names) + b"\0"
    return names, blob, (sample_id, start0, isize, flag, mapq, softclip)


@pytest.mark.gpu
def test_batcher_isolates_a_failing_payload(lib):
    """One bad window must not take its batch-mates down (mm_map never refuses, genotyper.cpp:387-393;
    an exception terminates Lancet2, core/async_worker.cpp:73-97): (i) a payload beyond the static caps
    throws in its own Enqueue; (ii) a payload that hits a device-side cap inside a shared batch is
    re-run alone and only its own Collect throws.  Every other payload's evidence equals the one-shot
    adapter's on the healthy groups."""
    from test_gpu_packed import homopolymer_group
    lib.lgr_adapter_isolation_dump.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + [C.c_void_p] * 6 + \
        [C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_longlong]
    lib.lgr_adapter_isolation_dump.restype = C.c_int
    rng = np.random.default_rng(23)
    ok = synth.make_groups(31, 5, n_reads=60, n_haps=3, hap_len=600)
    long_read = abi.Group(haps=[b"ACGT" * 200], reads=[b"ACGT" * 500], quals=[b"\x1e" * 2000], names=["n_long"], variants=[])
    hp = homopolymer_group(rng, 1200, 2, read_len=1000, flank=200)
    cap_hit = abi.Group(haps=hp.haps, reads=[hp.haps[0][195:1195]], quals=[bytes([30] * 1000)], names=["n_bad"], variants=hp.variants)

    def run(groups, mid_occ):
        batch = abi.Batch(groups)
        names, blob, cols = _meta(np.random.default_rng(1), groups, batch)
        status = np.full(len(groups), -1, dtype=np.int32)
        ctr = np.zeros(2, dtype=np.uint64)
        buf = C.create_string_buffer(64 << 20)
        bi = batch.c_struct()
        n = lib.lgr_adapter_isolation_dump(0, C.byref(bi), blob, b"normal\0tumor\0", *[c.ctypes.data for c in cols], mid_occ,
                                           status.ctypes.data, ctr.ctypes.data, buf, len(buf))
        assert n >= 0, buf.value.decode()
        per_group = {}
        for line in buf.value.decode().splitlines():
            g, rest = line.split(" ", 1)
            per_group.setdefault(int(g[1:]), []).append(rest)
        return status.tolist(), per_group, ctr

    # (i) static cap: refused at Enqueue, alone
    st, got, _ = run([ok[0], long_read, ok[1], ok[2]], 0)
    assert st == [0, 1, 0, 0]
    _, want, _ = run([ok[0], ok[1], ok[2]], 0)
    assert got[0] == want[0] and got[2] == want[1] and got[3] == want[2] and 1 not in got
    # (ii) device-side cap (> 65,535 anchors under a huge mid_occ): the batch is a partial success, the offender is
    # re-run alone, fails again and throws from its own Collect; the rest is untouched
    st, got, ctr = run([ok[3], cap_hit, ok[4]], 1000000)
    assert st == [0, 2, 0] and int(ctr[0]) == 1 and int(ctr[1]) == 1, (st, ctr)  # one shared batch, one re-run
    _, want, _ = run([ok[3], ok[4]], 1000000)
    assert got[0] == want[0] and got[2] == want[1] and len(want[0]) > 0


@pytest.mark.gpu
def test_cfg4_four_samples_evidence_per_sample(lib):
    """BASELINE configs[3] shape (four samples, colored-graph multi-sample calling): reads arrive in
    ReadCollector order (tag kind, then sample name, then qname: core/read_collector.cpp:42-54) and
    AddToTable keys the evidence by sample name in first-seen order (genotyper.cpp:423-456,
    support_array.cpp:19-28).  Through the batcher (4 workers, 8 payloads in flight each) every
    (variant, sample, allele) evidence block must equal what the oracle's assignments give."""
    spec = synth.TILED["cfg4"]
    groups = synth.make_tile_groups("cfg4", 42, 0, ref_len=80_000)
    assert len(groups) >= 10 and all(set(g.sample) == {0, 1, 2, 3} for g in groups)
    batch = abi.Batch(groups)
    names, blob, cols = _meta(None, groups, batch)
    sample_id = np.asarray([s for g in groups for s in g.sample], dtype=np.int32)
    cols = (sample_id,) + cols[1:]
    snames = tuple(spec["sample_names"])
    sblob = b"\0".join(s.encode() for s in snames) + b"\0"
    bi = batch.c_struct()
    buf = C.create_string_buffer(128 << 20)
    counters = np.zeros(17, dtype=np.uint64)
    n = lib.lgr_adapter_batcher_dump(0, C.byref(bi), blob, sblob, *[c.ctypes.data for c in cols], 4, 1, 8, 1,
                                     counters.ctypes.data, buf, len(buf))
    assert n > 0, buf.value.decode()[:500]
    got = buf.value.decode().splitlines()
    want, _ = O.oracle_genotype(batch, O.default_params(), n_threads=8)
    expect = expected_evidence(lib, batch, groups, names, *cols, want, snames)
    assert len(got) == len(expect) and len(got) > 40
    assert got == expect
    assert {line.split(" ")[2] for line in got} == {"S" + s for s in snames}  # every sample has evidence
