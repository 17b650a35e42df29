"""Read sharding across the GPUs of one box: independent units, no exchange step.

post_analysis shards reads, pre_survey shards plasmid pairs, batch mode shards mixtures.  Reads are
length-binned (sorted by predicted DP cells ~ N_q * sum(N_ref)) and dealt round-robin so that every rank
gets the same mix of long and short reads; results are gathered on the host in the original read
order (the order of my_fastq.keys(), post_analysis_core.py:63-65).  torch.distributed is used only
for the host-side gather (gloo or nccl); there is no collective on the data path.
"""
from __future__ import annotations

import numpy as np


def deal(weights, world: int) -> list[np.ndarray]:
    """indices per rank: sort by weight descending, deal in a snake (0..w-1, w-1..0) so that the sums
    of weights per rank stay within one item of each other."""
    order = np.argsort(-np.asarray(weights, dtype=np.int64), kind="stable")
    ranks = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        r = pos % (2 * world)
        ranks[r if r < world else 2 * world - 1 - r].append(int(idx))
    return [np.array(sorted(x), dtype=np.int64) for x in ranks]


def shard_reads(read_lengths, world: int, rank: int) -> np.ndarray:
    return deal(read_lengths, world)[rank]


def gather_in_order(local_items: list, local_indices, n_total: int, group=None) -> list | None:
    """all ranks call; rank 0 receives the items of every rank placed at their original indices."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    payload = (list(map(int, local_indices)), local_items)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * n_total
    for idxs, items in gathered:
        for i, it in zip(idxs, items):
            out[i] = it
    assert all(x is not None for x in out), "a read was not assigned to any rank"
    return out
