"""CPU check of the kernels' per-lane scalar core (lancet2_b200/csrc/lgr_core.cuh compiled by
g++, tests/hostemu) against the oracle: every alignment field, CIGAR and allele assignment
must be bit-identical.  This validates the device control flow without a GPU; the GPU tests
(test_gpu_parity.py) then only have to show that the kernels run the same code correctly."""
import numpy as np
import pytest

import hostemu_lib as H
import oracle_lib as O
from compare import compare_results
from lancet2_b200 import abi, synth


def str_group(rng, read_len, hap_len, n_haps, n_reads, unit_len=(1, 6), n_rep=(10, 60)):
    """haplotypes that differ by tandem-repeat copy number; reads with errors, Ns and
    reverse-complemented reads (exercises high-occurrence seeds, > 64 anchors, multi-chain)"""
    left, right = synth._rand_bases(rng, hap_len // 2), synth._rand_bases(rng, hap_len // 2)
    u = synth._rand_bases(rng, int(rng.integers(*unit_len)))
    base = int(rng.integers(*n_rep))
    haps = []
    for h in range(n_haps):
        cn = max(1, base + int(rng.integers(-4, 5))) if h > 0 else base
        haps.append(np.concatenate([left, np.tile(u, cn), right]))
    reads, quals, names = [], [], []
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in range(n_reads):
        hp = haps[int(rng.integers(0, n_haps))]
        st = int(rng.integers(0, max(1, hp.size - read_len)))
        rd = hp[st:st + read_len].copy()
        if rng.random() < 0.3:
            e = rng.random(rd.size) < 0.02
            rd[e] = synth._rand_bases(rng, int(e.sum()))
        if rng.random() < 0.2:
            rd = np.frombuffer(rd.tobytes().translate(comp)[::-1], dtype=np.uint8).copy()
        if rng.random() < 0.2:
            nn = rng.random(rd.size) < 0.03
            rd[nn] = ord("N")
        reads.append(rd.tobytes())
        quals.append(bytes(rng.integers(2, 42, rd.size, dtype=np.uint8)))
        names.append(f"s{i}")
    variants = []
    for h in range(1, n_haps):
        row = [(-1, 0, -1)] * n_haps
        row[0] = (left.size - 1, 1 + u.size, 0)
        row[h] = (left.size - 1, 1, 1)
        variants.append(row)
    return abi.Group(haps=[h.tobytes() for h in haps], reads=reads, quals=quals, names=names, variants=variants)


CASES = {
    "micro150": lambda: synth.make_groups(11, 4, n_reads=96, n_haps=5, hap_len=800),
    "L250": lambda: synth.make_groups(3, 2, read_len=250, hap_len=1500, n_haps=8, n_reads=48),
    "L100H300": lambda: synth.make_groups(4, 3, read_len=100, hap_len=300, n_haps=2, n_reads=64),
    "L1000": lambda: synth.make_groups(5, 2, read_len=1000, hap_len=3000, n_haps=3, n_reads=16, sub_err=0.01,
                                       indel_err=0.002),
    "noisy": lambda: synth.make_groups(6, 3, read_len=150, hap_len=600, n_haps=4, n_reads=64, sub_err=0.05,
                                       indel_err=0.003, n_frac=0.01),
    "short40": lambda: synth.make_groups(8, 2, read_len=40, hap_len=200, n_haps=3, n_reads=32),
    "STR150": lambda: [str_group(np.random.default_rng(70 + i), 150, 800, 4, 64) for i in range(4)],
    "STR250": lambda: [str_group(np.random.default_rng(80 + i), 250, 1000, 4, 32, (1, 4), (30, 120)) for i in range(3)],
    "STR600": lambda: [str_group(np.random.default_rng(90 + i), 600, 1500, 3, 16, (2, 30), (5, 40)) for i in range(2)],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_scalar_core_matches_oracle(name):
    batch = abi.Batch(CASES[name]())
    prm = O.default_params()
    want, st = O.oracle_genotype(batch, prm, n_threads=4)
    rc, got, st2 = H.emu_genotype(batch, prm)
    assert rc == 0
    errs = compare_results(batch, want, got)
    assert not errs, "\n".join(errs[:20])
    assert (st.chain_evals, st.n_anchors, st.dp_cells_full) == (st2.chain_evals, st2.n_anchors, st2.dp_cells_full)
    assert st2.dp_cells <= st2.dp_cells_full
    assert H.selfcheck()[0] == 0  # closed forms of the warp kernels vs the scalar paths (LGR_CORE_SELFCHECK)


def test_fixed_mid_occ_and_group_override():
    groups = [str_group(np.random.default_rng(5), 150, 600, 3, 48, (1, 3), (40, 80))]
    prm = O.default_params()
    for mid in (0, 10, 50):
        prm.mid_occ = mid
        batch = abi.Batch(groups)
        want, _ = O.oracle_genotype(batch, prm)
        rc, got, _ = H.emu_genotype(batch, prm)
        assert rc == 0 and not compare_results(batch, want, got)
    groups[0].mid_occ = 25
    prm.mid_occ = 0
    batch = abi.Batch(groups)
    want, _ = O.oracle_genotype(batch, prm)
    rc, got, _ = H.emu_genotype(batch, prm)
    assert rc == 0 and not compare_results(batch, want, got)


@pytest.mark.parametrize("k,w", [(15, 10), (11, 3), (13, 5)])
def test_other_minimizer_parameters(k, w):
    """generic ring-buffer sketch (w != 5) and the register-window sketch (w == 5) vs the oracle"""
    batch = abi.Batch(synth.make_groups(21, 2, n_reads=48, n_haps=3, hap_len=700, n_frac=0.01))
    prm = O.default_params()
    prm.k, prm.w = k, w
    want, _ = O.oracle_genotype(batch, prm)
    rc, got, _ = H.emu_genotype(batch, prm)
    assert rc == 0 and not compare_results(batch, want, got)
