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
