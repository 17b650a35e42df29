"""engine.SegmentedPipelineResult: the batches of one align_all call behave like one flat PipelineResult."""
import numpy as np


def _part(rng, n, stats_val):
    from multiplexnanopore_b200.engine import PipelineResult
    lens = rng.integers(0, 5, size=n)
    off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum(lens)
    rle = ((rng.integers(1, 50, size=int(off[-1])).astype(np.uint32) << 4) | rng.choice([1, 2, 4, 5, 7, 8], size=int(off[-1])).astype(np.uint32))
    return PipelineResult(rng.integers(-5, 99, size=n).astype(np.int32), rng.integers(0, 9, size=n).astype(np.int32),
                          (lens == 0).astype(np.uint8), off, rle, {"dp_cells": stats_val, "t_total": 1.0})


def test_segmented_result_equals_flat_concatenation():
    from multiplexnanopore_b200.engine import PipelineResult, SegmentedPipelineResult
    rng = np.random.default_rng(5)
    parts = [_part(rng, n, 10.0 * (k + 1)) for k, n in enumerate((12, 24, 1, 36))]
    seg = SegmentedPipelineResult(parts)
    offs = [parts[0].cigar_off]; base = int(parts[0].cigar_off[-1])
    for p in parts[1:]:
        offs.append(p.cigar_off[1:] + base); base += int(p.cigar_off[-1])
    flat = PipelineResult(np.concatenate([p.score for p in parts]), np.concatenate([p.new_query_seq_offset for p in parts]),
                          np.concatenate([p.status for p in parts]), np.concatenate(offs), np.concatenate([p.cigar_rle for p in parts]), {})
    assert np.array_equal(seg.score, flat.score) and np.array_equal(seg.status, flat.status)
    assert np.array_equal(seg.cigar_off, flat.cigar_off)
    for i in range(len(flat.score)):
        assert np.array_equal(seg.runs(i), flat.runs(i))
        assert seg.cigar_string(i) == flat.cigar_string(i) and seg.my_cigar(i) == flat.my_cigar(i) and seg.offset(i) == flat.offset(i)
    assert seg.nbytes == flat.nbytes
    assert np.array_equal(seg.cigar_rle, flat.cigar_rle)   # lazily concatenated
    assert seg.stats["dp_cells"] == 100.0 and seg.stats["t_total"] == 4.0


def test_batch_bounds_cover_all_queries_in_order():
    """savemoney._batch_bounds: ramped batch sizes must partition [0, n) whatever n, batch size and overlap are."""
    from multiplexnanopore_b200.savemoney import _batch_bounds
    for n in (1, 2, 999, 1000, 1001, 3999, 4000, 4001, 5499, 5500, 5501, 20000, 123457):
        for batch, overlap in ((1000, 6), (1000, 2), (1000, 8), (3, 6), (250, 4), (1, 6)):
            b = _batch_bounds(n, batch, overlap)
            assert b[0] == 0 and b[-1] == n and all(y > x for x, y in zip(b, b[1:])), (n, batch, overlap)
            assert max(y - x for x, y in zip(b, b[1:])) <= max(batch, 1), (n, batch, overlap)
    sizes = [y - x for x, y in zip(_batch_bounds(20000, 1000, 6), _batch_bounds(20000, 1000, 6)[1:])]
    assert sizes[:6] == [250, 250, 250, 500, 500, 500] and sizes[-6:] == [500, 500, 500, 250, 250, 250]
