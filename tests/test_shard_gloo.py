"""N > 1 host logic on CPU: world_size-2 gloo processes shard reads and gather results in read order."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from multiplexnanopore_b200 import shard


def test_deal_partitions_and_balances():
    rng = np.random.default_rng(0)
    w = rng.integers(100, 20000, size=1001)
    for world in (1, 2, 4, 8):
        parts = shard.deal(w, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(w)))
        sums = [int(w[p].sum()) for p in parts]
        assert max(sums) - min(sums) <= int(w.max())


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = np.random.default_rng(1).integers(50, 9000, size=n)
    mine = shard.shard_reads(lengths, world, rank)
    # stand-in for the per-rank engine call: any function of the read alone
    local = [(int(i), int(lengths[i]) * 3 + 1) for i in mine]
    out = shard.gather_in_order(local, mine, n)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_keeps_read_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n = 257
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lengths = np.random.default_rng(1).integers(50, 9000, size=n)
    assert out == [(i, int(lengths[i]) * 3 + 1) for i in range(n)]
