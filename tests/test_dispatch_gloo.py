"""world_size-2 (gloo, CPU) test of the multi-GPU host logic: groups are partitioned over
ranks, every rank processes only its shard (the oracle stands in for the GPU here — tests may
use it), results are gathered and put back in submission order; the merged result must equal
the single-process one bit for bit.  No collective touches the data path itself."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _digest(batch, res):
    return (res.aln[:batch.n_pairs].tobytes(), res.assign[:batch.n_assign].tobytes(),
            tuple(tuple(res.cigar(i)) for i in range(batch.n_pairs)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    from lancet2_b200 import abi, synth
    from lancet2_b200.dispatch import merge_in_submission_order, partition_groups
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    groups = synth.make_groups(5, 7, n_reads=24, n_haps=3, hap_len=500) + synth.make_region_groups(3, ref_len=20_000)[:3]
    shards = partition_groups(groups, world)
    mine = []
    for gi in shards[rank]:
        b = abi.Batch([groups[gi]])
        r, _ = O.oracle_genotype(b)
        mine.append(_digest(b, r))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = merge_in_submission_order(shards, gathered)
        single = []
        for g in groups:
            b = abi.Batch([g])
            r, _ = O.oracle_genotype(b)
            single.append(_digest(b, r))
        q.put((merged == single, [len(s) for s in shards], sorted(i for s in shards for i in s) == list(range(len(groups)))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_is_lossless_and_ordered():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, sizes, complete = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and complete and min(sizes) >= 1


def test_partition_balances_cost():
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from lancet2_b200 import synth
    from lancet2_b200.dispatch import group_cost, partition_groups
    rng = np.random.default_rng(1)
    groups = [synth.make_group(rng, n_reads=int(rng.integers(8, 64)), n_haps=int(rng.integers(2, 6)), hap_len=400) for _ in range(40)]
    for world in (1, 2, 4, 8):
        shards = partition_groups(groups, world)
        loads = [sum(group_cost(groups[i]) for i in s) for s in shards]
        assert sorted(i for s in shards for i in s) == list(range(len(groups)))
        assert max(loads) <= 1.25 * (sum(loads) / world) + max(group_cost(g) for g in groups)
    with pytest.raises(ValueError):
        partition_groups(groups, 0)


def _format_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import format_lib as F
    from lancet2_b200.dispatch import merge_in_submission_order, partition_supports
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(31)
    sups = [F.random_support(rng) for _ in range(60)] + [F.random_support(rng, n=300, n_alleles=2)]
    shards = partition_supports([len(s["allele"]) for s in sups], world)
    _, mine = F.emu_format([sups[i] for i in shards[rank]])   # the host build of the device core stands in for the GPU
    gathered = [None] * world
    dist.all_gather_object(gathered, [m.tobytes() for m in mine])
    if rank == 0:
        merged = merge_in_submission_order(shards, gathered)
        _, single = F.emu_format(sups)
        q.put((merged == [s.tobytes() for s in single], [len(s) for s in shards]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_of_supports():
    """the FORMAT-math row shards by support: records computed on two ranks and merged
    back equal the single-process records byte for byte"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_format_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and min(sizes) >= 1 and sum(sizes) == 61


def test_split_partition_is_exact_and_covers_every_group():
    """dispatch.partition_units_split (bench.py's strong-scaling shards at N > 1): every rank carries total / world of
    the cost, every unit's fractions tile [0, 1) without gap or overlap — also after rounding to a unit's groups —
    and a rank touches at most two partial units."""
    from lancet2_b200.dispatch import partition_units_split
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 4, 7, 8):
        for costs in ([10] * 50, rng.integers(1, 1000, 50).tolist(), [5], [3, 0, 9, 1]):
            shards = partition_units_split(costs, world)
            total = sum(costs)
            for s in shards:
                load = sum(costs[t] * (hi - lo) for t, lo, hi in s)
                assert abs(load - total / world) <= 1e-6 * total
                assert sum(1 for _, lo, hi in s if lo > 0 or hi < 1) <= 2
                assert [t for t, _, _ in s] == sorted(t for t, _, _ in s)
            for t, c in enumerate(costs):
                parts = sorted((lo, hi) for s in shards for tt, lo, hi in s if tt == t)
                if c == 0:
                    assert not parts
                    continue
                assert parts[0][0] == 0 and abs(parts[-1][1] - 1) < 1e-12
                n_groups = 237  # what bench.py does with a tile's groups
                taken = []
                for lo, hi in parts:
                    taken += list(range(int(round(lo * n_groups)), int(round(hi * n_groups))))
                assert taken == list(range(n_groups))
