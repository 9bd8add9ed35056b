"""GPU parity tests (run on the B200 box): the CUDA path, called through the C-ABI
(lgr_genotype_batch), against the CPU oracle on the same seeded inputs — bit-exact for every
alignment field, CIGAR, NM and per-(read, variant) allele assignment (f64 compared by bits)."""
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


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_oracle(gpu, name):
    batch = abi.Batch(CASES[name]())
    want, st = O.oracle_genotype(batch, gpu.params, n_threads=8)
    got, st2 = gpu.genotype_batch(batch)
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])
    assert (st.chain_evals, st.n_anchors, st.dp_cells_full, st.n_aligned) == \
        (st2.chain_evals, st2.n_anchors, st2.dp_cells_full, st2.n_aligned)
    assert st2.kernel_launches >= 8


def test_gpu_larger_microbench_batch(gpu):
    groups = synth.make_groups(137, 48, read_len=150, hap_len=1000, n_haps=8, n_reads=256)
    batch = abi.Batch(groups)
    want, _ = O.oracle_genotype(batch, gpu.params, n_threads=8)
    got, st = gpu.genotype_batch(batch)
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])


def test_resident_path_equals_batch_path_and_is_idempotent(gpu):
    batch = abi.Batch(synth.make_groups(271, 8, n_reads=128, n_haps=4, hap_len=700))
    a, _ = gpu.genotype_batch(batch)
    gpu.upload(batch)
    gpu.run_resident()
    gpu.run_resident()
    b = gpu.download(batch)
    assert not compare_results(batch, a, b)


def test_mid_occ_latch_matches_oracle(gpu):
    rng = np.random.default_rng(9)
    for _ in range(4):
        g = str_group(rng, 150, 600, 2, 4, (1, 3), (40, 200))
        for hap in g.haps:
            assert gpu.hap_mid_occ(hap) == O.load_oracle().orc_hap_mid_occ(gpu.params, hap, len(hap))


def test_empty_and_ragged_inputs(gpu):
    rng = np.random.default_rng(3)
    hap = synth._rand_bases(rng, 400).tobytes()
    g1 = abi.Group(haps=[hap], reads=[], quals=[], names=[], variants=[])          # no reads
    g2 = abi.Group(haps=[hap, hap[:200]], reads=[hap[10:160], b"ACGT", hap[100:101], b"N" * 50],
                   quals=[b"\x25" * 150, b"\x25" * 4, b"\x25", b"\x02" * 50], names=["a", "b", "c", "d"],
                   variants=[[(50, 1, 0), (50, 1, 1)]])
    g3 = abi.Group(haps=[hap], reads=[hap[0:150]], quals=[b"\x25" * 150], names=["e"], variants=[])  # no variants
    batch = abi.Batch([g1, g2, g3])
    want, _ = O.oracle_genotype(batch, gpu.params)
    got, _ = gpu.genotype_batch(batch)
    assert not compare_results(batch, want, got)
    empty = abi.Batch([])
    got, st = gpu.genotype_batch(empty)
    assert st.n_pairs == 0


def test_limits_fail_loudly(gpu):
    from lancet2_b200.realign import LgrError
    hap = b"A" * 500
    g = abi.Group(haps=[hap], reads=[b"C" * 2000], quals=[b"\x25" * 2000], names=["x"], variants=[])
    with pytest.raises(LgrError):
        gpu.genotype_batch(abi.Batch([g]))
    # haplotype + read beyond the ksw2 band of the option set: refused, not silently mis-aligned
    big = synth._rand_bases(np.random.default_rng(1), abi.LGR_MAX_HAP_LEN).tobytes()
    g = abi.Group(haps=[big], reads=[big[100:250]], quals=[b"\x25" * 150], names=["y"], variants=[])
    with pytest.raises(LgrError) as ei:
        gpu.genotype_batch(abi.Batch([g]))
    assert ei.value.code == -4


def test_submit_wait_tickets_overlap_and_match_oracle(gpu):
    """lgr_submit/lgr_wait: several batches in flight on one ctx, waited out of order, each
    bit-identical to the oracle; the slot limit and bad tickets fail loudly."""
    from lancet2_b200.realign import LgrError
    batches = [abi.Batch(synth.make_groups(900 + i, 3 + i, n_reads=64 + 16 * i, n_haps=2 + i, hap_len=500 + 100 * i))
               for i in range(abi.LGR_MAX_INFLIGHT)]
    for rnd in range(2):  # second round reuses the slots' device buffers
        tickets = [gpu.submit(b) for b in batches]
        assert sorted(t for t, _ in tickets) == list(range(abi.LGR_MAX_INFLIGHT))
        with pytest.raises(LgrError) as ei:
            gpu.submit(batches[0])
        assert ei.value.code == -7
        for (t, res), b in reversed(list(zip(tickets, batches))):
            st = gpu.wait(t)
            want, wst = O.oracle_genotype(b, gpu.params, n_threads=8)
            errs = compare_results(b, want, res)
            assert not errs, "\n".join(errs[:20])
            assert st.n_pairs == b.n_pairs and st.n_aligned == wst.n_aligned and st.h2d_bytes > 0 and st.d2h_bytes > 0
    with pytest.raises(LgrError):
        gpu.wait(0)  # nothing outstanding
    # assignments-only submission (what the Genotyper adapter asks for)
    t, res = gpu.submit(batches[1], want_aln=False)
    gpu.wait(t)
    want, _ = O.oracle_genotype(batches[1], gpu.params, n_threads=8)
    assert not compare_results(batches[1], want, res, check_aln=False)


def test_full_cfg2_workload_matches_oracle(gpu):
    """The whole bench workload (BASELINE cfg2: 1 Mb region, 30x/30x, ~233 K pairs) against the
    oracle, every field, not a sample."""
    batch = abi.Batch(synth.make_region_groups(42, ref_len=1_000_000))
    assert batch.n_pairs > 200_000
    want, wst = O.oracle_genotype(batch, gpu.params, n_threads=16)
    got, st = gpu.genotype_batch(batch)
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])
    assert (st.n_aligned, st.chain_evals, st.n_anchors, st.dp_cells_full) == \
        (wst.n_aligned, wst.chain_evals, wst.n_anchors, wst.dp_cells_full)


def test_size_independent_properties_at_microbench_size(gpu):
    """cfg5-sized batch (2 M pairs): properties that need no oracle.
    (i) a read that is an exact substring of a haplotype aligns to it as one M run, NM 0, at
    the position it was cut from (unique 150-mers in random sequence); (ii) results do not
    depend on where a group sits in the batch; (iii) the device is deterministic."""
    rng = np.random.default_rng(77)
    groups, truth = [], []
    for gi in range(1024):
        g = synth.make_group(rng, read_len=150, hap_len=1000, n_haps=8, n_reads=256, name_prefix=f"g{gi}_")
        reads, pos = list(g.reads), []
        for r in range(0, 256, 4):  # every 4th read: exact cut from a known haplotype
            h = int(rng.integers(0, 8))
            st0 = int(rng.integers(0, len(g.haps[h]) - 150))
            reads[r] = g.haps[h][st0:st0 + 150]
            pos.append((r, h, st0))
        groups.append(abi.Group(haps=g.haps, reads=reads, quals=g.quals, names=g.names, variants=g.variants))
        truth.append(pos)
    batch = abi.Batch(groups)
    assert batch.n_pairs == 1024 * 256 * 8
    got, st = gpu.genotype_batch(batch)
    m_op = 150 << 4
    for gi, pos in enumerate(truth):
        r0 = batch.grp_read_begin[gi]
        for r, h, st0 in pos:
            pair = int(batch.pair_off[r0 + r]) + h
            a = got.aln[pair]
            assert a["valid"] == 1 and a["nm"] == 0 and a["qs"] == 0 and a["qe"] == 150, (gi, r, h)
            assert a["rs"] == st0 and a["re"] == st0 + 150 and got.cigar(pair) == [m_op], (gi, r, h)
    # (ii) + (iii): reversed group order gives the same per-group records
    rev = abi.Batch(groups[::-1])
    got2, _ = gpu.genotype_batch(rev)
    got3, _ = gpu.genotype_batch(batch)
    # (cigar_off is an arena offset handed out by an atomic: the one field that may differ between runs)
    for f in got.aln.dtype.names:
        if f != "cigar_off":
            assert np.array_equal(got3.aln[f], got.aln[f]), f
    for pair in np.nonzero(got.aln["cigar_off"] >= 0)[0][:2000]:
        assert got.cigar(int(pair)) == got3.cigar(int(pair))
    assert got3.assign.tobytes() == got.assign.tobytes()
    G = len(groups)
    for gi in (0, 1, 17, 511, G - 1):
        p0, p1 = int(batch.pair_off[batch.grp_read_begin[gi]]), int(batch.pair_off[batch.grp_read_begin[gi + 1]])
        q0, q1 = int(rev.pair_off[rev.grp_read_begin[G - 1 - gi]]), int(rev.pair_off[rev.grp_read_begin[G - gi]])
        fields = ["valid", "score", "rs", "re", "qs", "qe", "rev", "dp_score", "dp_max", "mlen", "blen", "nm", "n_cigar"]
        for f in fields:
            assert np.array_equal(got.aln[f][p0:p1], got2.aln[f][q0:q1]), (gi, f)
        a0, a1 = int(batch.asg_off[batch.grp_read_begin[gi]]), int(batch.asg_off[batch.grp_read_begin[gi + 1]])
        b0, b1 = int(rev.asg_off[rev.grp_read_begin[G - 1 - gi]]), int(rev.asg_off[rev.grp_read_begin[G - gi]])
        assert got.assign[a0:a1].tobytes() == got2.assign[b0:b1].tobytes()


def test_alphabet_quality_and_size_edges(gpu):
    """lower case, IUPAC codes, U, Phred 0 / 93 / 255, a read shorter than k, an all-N read,
    a read of exactly LGR_MAX_READ_LEN on the longest haplotype the ksw2 band admits, 64 haplotypes
    in one group, a variant no haplotype carries and a zero-length allele."""
    rng = np.random.default_rng(123)
    hap = synth._rand_bases(rng, 900).tobytes()
    low = hap[100:250].lower()
    iupac = bytearray(hap[300:450]); iupac[20] = ord("R"); iupac[70] = ord("y"); iupac[100] = ord("U")
    quals = [bytes([0] * 150), bytes([93] * 150), bytes([255] * 150), bytes(rng.integers(0, 94, 150, dtype=np.uint8))]
    g1 = abi.Group(haps=[hap, hap[:400] + b"ACG" + hap[400:]],
                   reads=[low, bytes(iupac), hap[500:650], hap[395:545], b"ACGTA", b"N" * 150],
                   quals=quals + [b"\x1e" * 5, b"\x02" * 150], names=[f"e{i}" for i in range(6)],
                   variants=[[(400, 0, 0), (400, 3, 1)], [(-1, 0, -1), (-1, 0, -1)], [(10, 5, 0), (10, 5, 0)]])
    # with the reference's bw = 10000 the ksw2 band (1.5 bw + 1) must cover haplotype + read, which
    # the device path insists on: 13,900 + 1,024 is just inside (windows are <= 2,500 bp upstream)
    big_hap = synth._rand_bases(rng, 13_900).tobytes()
    long_read = bytearray(big_hap[3000:3000 + abi.LGR_MAX_READ_LEN]); long_read[500] = ord("A") if long_read[500] != ord("A") else ord("C")
    g2 = abi.Group(haps=[big_hap], reads=[bytes(long_read), big_hap[100:250]],
                   quals=[b"\x25" * abi.LGR_MAX_READ_LEN, b"\x25" * 150], names=["big", "small"], variants=[[(3500, 1, 0)]])
    g3 = synth.make_group(rng, read_len=150, hap_len=600, n_haps=64, n_reads=32)
    batch = abi.Batch([g1, g2, g3])
    want, _ = O.oracle_genotype(batch, gpu.params, n_threads=8)
    got, _ = gpu.genotype_batch(batch)
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])
    assert got.aln[0]["valid"] == 1 and got.aln[0]["nm"] == 0      # lower case read == upper case haplotype
    assert got.aln[int(batch.pair_off[4])]["valid"] == 0            # 5 bp read: no minimizer


def _params(**kw):
    p = O.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


OPTION_SETS = {
    # other minimizer geometry: even k takes the lane-per-haplotype sketch (symmetric k-mers are
    # skipped), w != 5 the ring-buffer sketch, k = 15 the 64-bit window
    "k12w5": dict(k=12, w=5),
    "k15w10": dict(k=15, w=10),
    "k9w3": dict(k=9, w=3),
    # other scoring: minimap2's own sr preset values and a cheap-gap set
    "sr_preset_scores": dict(a=2, b=8, q=12, e=2, end_bonus=20000),
    "cheap_gaps": dict(a=1, b=2, q=3, e=1),
    # chaining knobs
    "chain_knobs": dict(max_chain_skip=5, max_chain_iter=50, min_cnt=2, min_chain_score=25, max_gap=100),
    "fixed_mid_occ": dict(mid_occ=3),
}


@pytest.mark.parametrize("name", sorted(OPTION_SETS))
def test_other_option_sets_match_oracle(name):
    """the kernels are not specialised to the reference's option values: every set must stay
    bit-identical to the oracle run with the same lgr_params (or be refused by validate_params)"""
    from lancet2_b200.realign import GpuRealigner, LgrError
    prm = _params(**OPTION_SETS[name])
    rng = np.random.default_rng(31)
    groups = synth.make_groups(41, 4, n_reads=96, n_haps=4, hap_len=700) + \
        synth.make_groups(42, 2, read_len=250, hap_len=1200, n_haps=3, n_reads=48, sub_err=0.02, indel_err=0.002) + \
        [str_group(rng, 150, 700, 3, 48)]
    batch = abi.Batch(groups)
    try:
        g = GpuRealigner(0, params=prm)
    except LgrError as e:
        pytest.skip(f"option set refused by the device path: {e}")
    try:
        want, wst = O.oracle_genotype(batch, prm, n_threads=8)
        got, st = g.genotype_batch(batch)
        errs = compare_results(batch, want, got)
        assert not errs, "\n".join(errs[:20])
        assert (st.n_aligned, st.chain_evals, st.n_anchors) == (wst.n_aligned, wst.chain_evals, wst.n_anchors)
        assert st.n_aligned > 0
    finally:
        g.close()


def test_full_microbench_point_matches_oracle(gpu):
    """the cfg5 bench point at full size (2.1 M pairs) against the oracle, every field"""
    batch = abi.Batch(synth.make_groups(42, 1024, read_len=150, hap_len=1000, n_haps=8, n_reads=256))
    assert batch.n_pairs == 2_097_152
    want, wst = O.oracle_genotype(batch, gpu.params, n_threads=24)
    got, st = gpu.genotype_batch(batch, arena=1 << 22)
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])
    assert (st.n_aligned, st.chain_evals, st.dp_cells_full) == (wst.n_aligned, wst.chain_evals, wst.dp_cells_full)
