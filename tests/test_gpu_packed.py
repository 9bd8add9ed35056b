"""GPU tests of the packed wire format, the host-driven overflow pass and the per-group status
(run on the B200 box through the C-ABI): the packed path must produce byte-identical results to
the plain path, nothing mm_map accepts may be refused, and one pathological group must not fail
the others."""
import numpy as np
import pytest

import oracle_lib as O
from compare import compare_results
from lancet2_b200 import abi, synth
from test_hostemu_parity import CASES, str_group

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from lancet2_b200.realign import GpuRealigner
    g = GpuRealigner(0)
    yield g
    g.close()


def alphabet_group():
    rng = np.random.default_rng(5)
    hap = synth._rand_bases(rng, 777).tobytes()
    r1 = bytearray(hap[100:250]); r1[7] = ord("N"); r1[99] = ord("r"); r1[33] = ord("U")
    reads = [hap[10:160].lower(), bytes(r1), b"N" * 40, hap[600:777], b"ACG"]
    quals = [bytes(rng.integers(0, 94, len(r), dtype=np.uint8)) for r in reads]
    hap2 = bytearray(hap); hap2[300] = ord("n")
    return abi.Group(haps=[hap, bytes(hap2)], reads=reads, quals=quals, names=[f"a{i}" for i in range(len(reads))],
                     variants=[[(300, 1, 0), (300, 1, 1)]])


def same_bytes(batch, a, b):
    errs = compare_results(batch, a, b)
    assert not errs, "\n".join(errs[:20])
    assert a.assign[:batch.n_assign].tobytes() == b.assign[:batch.n_assign].tobytes()


@pytest.mark.parametrize("name", sorted(CASES) + ["alphabet"])
def test_packed_path_matches_plain_path(gpu, name):
    groups = [alphabet_group()] + synth.make_groups(2, 2, n_reads=33, n_haps=3, hap_len=431) if name == "alphabet" else CASES[name]()
    batch = abi.Batch(groups)
    packed = abi.PackedBatch(groups, gpu.lib)
    plain, st1 = gpu.genotype_batch(batch)
    got, st2 = gpu.genotype_packed(packed, batch)
    same_bytes(batch, plain, got)
    assert (st1.n_aligned, st1.chain_evals, st1.n_anchors, st1.dp_cells_full) == (st2.n_aligned, st2.chain_evals, st2.n_anchors, st2.dp_cells_full)
    assert st2.h2d_bytes == packed.slab_bytes  # ONE copy: the slab with its directory
    assert np.array_equal(plain.grp_mid_occ[:batch.n_groups], got.grp_mid_occ[:batch.n_groups])
    want, _ = O.oracle_genotype(batch, gpu.params, n_threads=8)
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])


def test_packed_submit_resident_and_latched_mid_occ(gpu):
    groups = synth.make_groups(19, 6, n_reads=80, n_haps=4, hap_len=650)
    groups[2].mid_occ = 23  # a worker's latched value travels in the directory
    batch = abi.Batch(groups)
    packed = abi.PackedBatch(groups, gpu.lib)
    want, _ = O.oracle_genotype(batch, gpu.params, n_threads=8)
    tickets = [gpu.submit_packed(packed, batch) for _ in range(3)]
    for t, res in tickets:
        gpu.wait(t)
        errs = compare_results(batch, want, res)
        assert not errs, "\n".join(errs[:20])
        assert res.grp_mid_occ[2] == 23 and res.grp_mid_occ[0] == 10
    gpu.upload_packed(packed)
    gpu.run_resident()
    st = gpu.run_resident()
    res = gpu.download(batch)
    assert not compare_results(batch, want, res)
    assert st.kernel_launches >= 10
    # an empty packed batch and a batch of empty groups
    e = abi.PackedBatch([], gpu.lib)
    _, st = gpu.genotype_packed(e, abi.Batch([]))
    assert st.n_pairs == 0
    g0 = [abi.Group(haps=[b"ACGTACGTAC"], reads=[], quals=[], names=[], variants=[])]
    _, st = gpu.genotype_packed(abi.PackedBatch(g0, gpu.lib), abi.Batch(g0))
    assert st.n_pairs == 0


def homopolymer_group(rng, run, n_reads, read_len=150, flank=300):
    """reads inside a homopolymer / dinucleotide run: every read minimizer occurs ~run times on the
    haplotype, so with a large mid_occ a pair carries (read minimizers x run) anchors"""
    left, right = synth._rand_bases(rng, flank), synth._rand_bases(rng, flank)
    unit = synth._rand_bases(rng, 2)
    while unit[0] == unit[1]:
        unit = synth._rand_bases(rng, 2)
    rep = np.tile(unit, run // 2)
    hap0 = np.concatenate([left, rep, right])
    hap1 = np.concatenate([left, rep[:-4], right])
    reads, quals, names = [], [], []
    for i in range(n_reads):
        st = flank - 10 + int(rng.integers(0, 20)) if i % 2 == 0 else int(rng.integers(0, hap0.size - read_len))
        rd = hap0[st:st + read_len]
        reads.append(rd.tobytes()), quals.append(bytes([30] * rd.size)), names.append(f"h{i}")
    return abi.Group(haps=[hap0.tobytes(), hap1.tobytes()], reads=reads, quals=quals, names=names,
                     variants=[[(flank + rep.size - 5, 5, 0), (flank + rep.size - 5, 1, 1)]])


def test_no_refusal_beyond_16384_anchors():
    """mm_map never refuses a (read, haplotype) pair (genotyper.cpp:387-393).  With a large fixed
    mid_occ the reads inside the repeat carry > 16,384 anchors (round 1 refused such a batch): they
    take the overflow pass, whose workspace is sized from what the batch needs, and match the oracle."""
    from lancet2_b200.realign import GpuRealigner
    prm = O.default_params()
    prm.mid_occ = 100000
    rng = np.random.default_rng(404)
    groups = [homopolymer_group(rng, 300, 6), synth.make_group(rng, n_reads=40, n_haps=3, hap_len=500)]
    batch = abi.Batch(groups)
    want, wst = O.oracle_genotype(batch, prm, n_threads=8)
    assert wst.n_anchors / max(1, batch.n_pairs) > 1000 and wst.n_anchors > 3 * 16384
    g = GpuRealigner(0, params=prm)
    try:
        got, st = g.genotype_batch(batch)
        errs = compare_results(batch, want, got)
        assert not errs, "\n".join(errs[:20])
        assert st.reserved > 0  # pairs that went through the overflow pass
        assert (st.n_anchors, st.chain_evals, st.n_aligned) == (wst.n_anchors, wst.chain_evals, wst.n_aligned)
        # the same through the asynchronous call (the overflow pass runs inside lgr_wait)
        t, res = g.submit(batch)
        g.wait(t)
        assert not compare_results(batch, want, res)
    finally:
        g.close()


def test_one_pathological_group_does_not_fail_the_batch():
    """a pair beyond what the 16-bit chain workspace can index (> 65,535 anchors; needs a ~1 kb read
    inside a repeat AND mid_occ far above the reference's) fails only its own group: the call
    returns LGR_E_PARTIAL, grp_status names the group, every other group equals the oracle."""
    from lancet2_b200.realign import GpuRealigner, LgrError
    prm = O.default_params()
    prm.mid_occ = 1000000
    rng = np.random.default_rng(11)
    bad = homopolymer_group(rng, 1200, 2, read_len=1000, flank=200)
    bad = abi.Group(haps=bad.haps, reads=[bad.haps[0][195:1195]], quals=[bytes([30] * 1000)], names=["bad"], variants=bad.variants)
    ok1 = synth.make_group(rng, n_reads=50, n_haps=3, hap_len=600)
    ok2 = synth.make_group(rng, n_reads=30, n_haps=2, hap_len=400)
    batch = abi.Batch([ok1, bad, ok2])
    g = GpuRealigner(0, params=prm)
    try:
        with pytest.raises(LgrError) as ei:  # without the status array the whole call fails, as before
            g.genotype_batch(batch)
        assert ei.value.code == abi.LGR_E_LIMIT
        res, st = g.genotype_batch(batch, group_status=True)
        assert res.rc == abi.LGR_E_PARTIAL
        assert res.grp_status[:3].tolist() == [0, abi.LGR_E_LIMIT, 0]
        good = abi.Batch([ok1, ok2])
        want, _ = O.oracle_genotype(good, prm, n_threads=8)
        a1 = batch.n_reads and int(batch.asg_off[batch.grp_read_begin[1]])
        a2 = int(batch.asg_off[batch.grp_read_begin[2]])
        got_assign = np.concatenate([res.assign[:a1], res.assign[a2:batch.n_assign]])
        assert got_assign.tobytes() == want.assign[:good.n_assign].tobytes()
    finally:
        g.close()


def test_many_contexts_share_one_gpu():
    """one context per worker thread.  64 contexts, each running a Genotype()-sized
    batch, must fit comfortably: scratch scales with the batch, the overflow workspace is lazy."""
    import torch
    from lancet2_b200.realign import GpuRealigner
    groups = synth.make_groups(77, 1, n_reads=300, n_haps=3, hap_len=900)
    batch = abi.Batch(groups)
    free0, _ = torch.cuda.mem_get_info(0)
    ctxs = [GpuRealigner(0) for _ in range(64)]
    try:
        ref = None
        for c in ctxs:
            res, _ = c.genotype_batch(batch)
            if ref is None:
                ref = res
            else:
                assert res.assign[:batch.n_assign].tobytes() == ref.assign[:batch.n_assign].tobytes()
        free1, _ = torch.cuda.mem_get_info(0)
        assert (free0 - free1) / 64 < 64 * 2**20, f"{(free0 - free1) / 64 / 2**20:.1f} MiB per context"
    finally:
        for c in ctxs:
            c.close()


def test_reserve_presizes_slots_and_results_do_not_change():
    """lgr_reserve (what the batcher calls at construction): the submission slots and their arenas exist
    before the first batch, no arena grows while batches that fit are in flight, results are the same."""
    from lancet2_b200.realign import GpuRealigner
    groups = synth.make_region_groups(13, ref_len=40_000)
    batch = abi.Batch(groups)
    plain = GpuRealigner(0)
    g = GpuRealigner(0)
    try:
        ref, _ = plain.genotype_batch(batch)
        assert g.lib.lgr_arena_bytes(g._ctx) == 0
        assert g.lib.lgr_reserve(g._ctx, 256 << 20, 3) == 0
        reserved = g.lib.lgr_arena_bytes(g._ctx)
        assert reserved >= 3 * (256 << 20)
        packed = abi.PackedBatch(groups, g.lib)
        import torch
        packed.pin(torch)
        tickets = []
        outs = []
        for _ in range(3):
            t, res = g.submit_packed(packed, batch, want_aln=False)
            tickets.append(t), outs.append(res)
        for t in tickets:
            g.wait(t)
        assert g.lib.lgr_arena_bytes(g._ctx) == reserved          # nothing grew
        for res in outs:
            assert res.assign[:batch.n_assign].tobytes() == ref.assign[:batch.n_assign].tobytes()
        assert g.lib.lgr_reserve(g._ctx, 1 << 20, abi.LGR_MAX_INFLIGHT + 1) == -1
        assert g.lib.lgr_reserve(g._ctx, -1, 1) == -1
    finally:
        g.close()
        plain.close()
