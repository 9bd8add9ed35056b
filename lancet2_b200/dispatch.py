"""Host-side sharding of Genotype() payloads over the GPUs of one box.

Windows (and the graph components inside them) are independent in the reference
(docs/guides/architecture.md:124; one `Genotyper` per worker thread,
core/variant_builder.h:94), so the multi-GPU story is a partition of whole groups with NO
data-path collective: every rank owns one GPU, runs `lgr_genotype_batch` on its shard and the
results are put back in submission order on the host (what `VariantStore`'s coordinate sort
does in the reference, core/variant_store.cpp:101-122).
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

from .abi import Group


def group_cost(g: Group) -> int:
    """work estimate of one group: every read is mapped against every haplotype
    (O(H x R x L), docs/guides/architecture.md:99)"""
    hap = sum(len(h) for h in g.haps)
    return max(1, len(g.reads)) * max(1, hap)


def partition_groups(groups: Sequence[Group], world: int) -> List[List[int]]:
    """longest-processing-time greedy partition of group indices over `world` ranks;
    deterministic (ties by index), every rank's list is ascending."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(groups)), key=lambda i: (-group_cost(groups[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + group_cost(groups[i]), r))
    for s in shards:
        s.sort()
    return shards


def partition_units(costs: Sequence[int], world: int) -> List[List[int]]:
    """the same longest-processing-time partition for any independent units with given costs (the 1 Mb
    tiles of the sharded bench workloads, synth.TILED); deterministic, ascending per rank"""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + int(costs[i]), r))
    for s in shards:
        s.sort()
    return shards


def merge_in_submission_order(shards: Sequence[Sequence[int]], per_rank_results: Sequence[Sequence[object]]) -> List[object]:
    """inverse of partition_groups: results[i] is the result of group i"""
    n = sum(len(s) for s in shards)
    out: List[object] = [None] * n
    for idx, res in zip(shards, per_rank_results):
        if len(idx) != len(res):
            raise ValueError("a rank returned a different number of results than it was given groups")
        for i, r in zip(idx, res):
            out[i] = r
    return out


def support_cost(n_records: int) -> int:
    """work estimate of one VariantSupport in the FORMAT-math kernels: linear passes
    plus the O(n^2) rank / bin scans"""
    n = max(0, int(n_records))
    return 64 + 16 * n + n * n // 8


def partition_supports(n_records: Sequence[int], world: int) -> List[List[int]]:
    """the same longest-processing-time partition for supports (one variant x one sample each):
    supports are independent, so every rank runs `lgr_format_metrics` on its shard and the
    records are merged back with `merge_in_submission_order`; no collective."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(n_records)), key=lambda i: (-support_cost(n_records[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + support_cost(n_records[i]), r))
    for s in shards:
        s.sort()
    return shards
