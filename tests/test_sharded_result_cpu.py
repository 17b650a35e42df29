"""Row (e), host logic without a GPU: the view that puts the per-GPU results of a sharded alignment stage back into
the caller's order (ShardedPipelineResult) against a direct construction, for reads (2 * N_ref pair-strands per unit)
and for plasmid pairs (2 per unit)."""
import numpy as np
import pytest

from multiplexnanopore_b200 import shard
from multiplexnanopore_b200.engine import PipelineResult, SegmentedPipelineResult, ShardedPipelineResult


def fake_result(rng, n):
    lens = rng.integers(0, 5, size=n)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(lens)
    rle = ((rng.integers(1, 1 << 12, size=int(off[-1])) << 4) | rng.choice([1, 2, 4, 5, 7, 8], size=int(off[-1]))).astype(np.uint32)
    return PipelineResult(rng.integers(-50, 900, size=n).astype(np.int32), rng.integers(0, 9000, size=n).astype(np.int32),
                          rng.integers(0, 2, size=n).astype(np.uint8), off, rle, {"dp_cells": float(n), "kernel_launches": 3.0})


@pytest.mark.parametrize("world,per_unit,n_units", [(2, 12, 37), (3, 2, 50), (8, 12, 9), (4, 6, 4)])
def test_sharded_view_restores_unit_order(world, per_unit, n_units):
    rng = np.random.default_rng(world * 100 + per_unit)
    whole = fake_result(rng, n_units * per_unit)
    index = shard.deal(rng.integers(100, 9000, size=n_units), world)
    parts = []
    for ix in index:
        rows = (ix[:, None] * per_unit + np.arange(per_unit)[None, :]).ravel()
        lens = np.diff(whole.cigar_off)[rows]
        off = np.zeros(len(rows) + 1, np.int64)
        off[1:] = np.cumsum(lens)
        rle = np.concatenate([whole.runs(int(r)) for r in rows]) if len(rows) else np.empty(0, np.uint32)
        p = PipelineResult(whole.score[rows], whole.new_query_seq_offset[rows], whole.status[rows], off, rle.astype(np.uint32),
                           {"dp_cells": float(len(rows)), "kernel_launches": 3.0})
        if len(rows) > 3 * per_unit:  # some shards arrive as several batches
            cut = (len(ix) // 2) * per_unit
            a = PipelineResult(p.score[:cut], p.new_query_seq_offset[:cut], p.status[:cut], off[:cut + 1].copy(), rle[:off[cut]].astype(np.uint32), dict(p.stats))
            b = PipelineResult(p.score[cut:], p.new_query_seq_offset[cut:], p.status[cut:], off[cut:] - off[cut], rle[off[cut]:].astype(np.uint32), {k: 0.0 for k in p.stats})
            p = SegmentedPipelineResult([a, b])
        parts.append(p)
    keep = [k for k, ix in enumerate(index) if len(ix)]
    got = ShardedPipelineResult([parts[k] for k in keep], [index[k] for k in keep], per_unit)
    assert np.array_equal(got.score, whole.score)
    assert np.array_equal(got.new_query_seq_offset, whole.new_query_seq_offset)
    assert np.array_equal(got.status, whole.status)
    assert np.array_equal(got.cigar_off, whole.cigar_off)
    for i in range(n_units * per_unit):
        assert np.array_equal(got.runs(i), whole.runs(i))
        assert got.cigar_string(i) == whole.cigar_string(i)
    assert np.array_equal(got.cigar_rle, whole.cigar_rle)
    assert got.stats["dp_cells"] == n_units * per_unit
