"""Differential fuzz: randomly shaped groups (read length, haplotype length and count, error
rates, N bases, reverse-complement reads, low-complexity inserts, chimeric reads, random option
sets) through the oracle and through the kernels' per-lane core compiled for the host
(tests/hostemu) on the CPU, and — on the GPU box — through the CUDA path.  Every field must agree."""
import numpy as np
import pytest

import hostemu_lib as H
import oracle_lib as O
from compare import compare_results
from lancet2_b200 import abi, synth

COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def fuzz_group(rng) -> abi.Group:
    read_len = int(rng.choice([36, 75, 100, 150, 150, 150, 250, 400]))
    hap_len = int(rng.integers(max(read_len + 20, 120), 1600))
    n_haps = int(rng.integers(1, 7))
    ref = synth._rand_bases(rng, hap_len)
    if rng.random() < 0.3:  # low-complexity insert
        u = synth._rand_bases(rng, int(rng.integers(1, 7)))
        p = int(rng.integers(20, hap_len - 20))
        ref = np.concatenate([ref[:p], np.tile(u, int(rng.integers(5, 40))), ref[p:]])
    haps, variants = [ref], []
    for h in range(1, n_haps):
        alt, rb, ab = synth._spike_variant(rng, ref)
        haps.append(alt)
        row = [(-1, 0, -1)] * n_haps
        row[0], row[h] = (rb[0], rb[1], 0), (ab[0], ab[1], 1)
        variants.append(row)
    sub, ins = float(rng.choice([0.0, 0.002, 0.01, 0.05])), float(rng.choice([0.0, 0.0005, 0.005]))
    reads, quals, names = [], [], []
    for i in range(int(rng.integers(1, 24))):
        hp = haps[int(rng.integers(0, n_haps))]
        st = int(rng.integers(-20, max(1, hp.size - read_len + 20)))
        idx = np.arange(st, st + read_len)
        inside = (idx >= 0) & (idx < hp.size)
        rd = synth._rand_bases(rng, read_len)
        rd[inside] = hp[idx[inside]]
        e = rng.random(read_len) < sub
        rd[e] = synth._rand_bases(rng, int(e.sum()))
        if ins > 0 and rng.random() < ins * read_len and read_len > 20:
            p = int(rng.integers(5, read_len - 5))
            d = int(rng.integers(1, 12))
            rd = np.concatenate([rd[:p], rd[p + d:], synth._rand_bases(rng, d)]) if rng.random() < 0.5 else \
                np.concatenate([rd[:p], synth._rand_bases(rng, d), rd[p:-d]])
        if rng.random() < 0.1:  # chimeric: second half from elsewhere
            q = hp[int(rng.integers(0, max(1, hp.size - read_len))):][:rd.size // 2]
            rd[rd.size - q.size:] = q
        if rng.random() < 0.15:
            nn = rng.random(rd.size) < 0.02
            rd[nn] = ord("N")
        b = rd.tobytes()
        if rng.random() < 0.2:
            b = b.translate(COMP)[::-1]
        if rng.random() < 0.05:
            b = b.lower()
        reads.append(b)
        quals.append(bytes(rng.integers(0, 60, len(b), dtype=np.uint8)))
        names.append(f"f{i}_{int(rng.integers(0, 1 << 30))}")
    return abi.Group(haps=[h.tobytes() for h in haps], reads=reads, quals=quals, names=names, variants=variants)


def fuzz_params(rng):
    p = O.default_params()
    if rng.random() < 0.5:
        k, w = [(11, 5), (13, 5), (9, 5), (15, 10), (12, 5), (11, 3)][int(rng.integers(0, 6))]
        p.k, p.w = k, w
    if rng.random() < 0.3:
        p.a, p.b, p.q, p.e = [(1, 4, 12, 3), (1, 2, 3, 1), (1, 4, 6, 2), (2, 8, 12, 2)][int(rng.integers(0, 4))]
        p.end_bonus = 20000
    if rng.random() < 0.3:
        p.min_cnt, p.min_chain_score = int(rng.integers(1, 4)), int(rng.integers(15, 45))
    if rng.random() < 0.2:
        p.mid_occ = int(rng.integers(2, 30))
    return p


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_scalar_core_vs_oracle(seed):
    rng = np.random.default_rng(5000 + seed)
    for _ in range(6):
        prm = fuzz_params(rng)
        batch = abi.Batch([fuzz_group(rng) for _ in range(int(rng.integers(1, 5)))])
        want, st = O.oracle_genotype(batch, prm, n_threads=2)
        rc, got, st2 = H.emu_genotype(batch, prm)
        assert rc == 0
        errs = compare_results(batch, want, got)
        assert not errs, "\n".join(errs[:20])
        assert (st.chain_evals, st.n_anchors, st.dp_cells_full) == (st2.chain_evals, st2.n_anchors, st2.dp_cells_full)
    # the closed forms the warp kernels use (co-linear chain, exact-match / overhang extension) were
    # checked against the scalar loops on every pair whose precondition held
    failures, n_colinear, n_ext, n_tail, _ = H.selfcheck()
    assert failures == 0 and n_colinear > 0 and n_ext > 0 and n_tail > 0, (failures, n_colinear, n_ext, n_tail)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(16))
def test_fuzz_gpu_vs_oracle(seed):
    from lancet2_b200.realign import GpuRealigner, LgrError
    rng = np.random.default_rng(9000 + seed)
    for _ in range(5):
        prm = fuzz_params(rng)
        batch = abi.Batch([fuzz_group(rng) for _ in range(int(rng.integers(4, 24)))])
        try:
            g = GpuRealigner(0, params=prm)
        except LgrError:
            continue  # option set outside what validate_params admits
        try:
            want, st = O.oracle_genotype(batch, prm, n_threads=8)
            try:
                got, st2 = g.genotype_batch(batch)
            except LgrError as e:
                assert e.code == -4  # a pair beyond a device cap (e.g. > 16384 anchors in a long repeat): refused loudly
                continue
            errs = compare_results(batch, want, got)
            assert not errs, "\n".join(errs[:20])
            for f in ("n_aligned", "chain_evals", "n_anchors", "dp_cells_full"):
                assert getattr(st, f) == getattr(st2, f), (f, getattr(st, f), getattr(st2, f), prm.k, prm.w, prm.mid_occ)
        finally:
            g.close()
