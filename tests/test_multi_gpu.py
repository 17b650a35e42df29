"""Row (e) on the device: reads (post_analysis) and plasmid pairs (pre_survey) sharded over two GPUs give exactly the
result of one GPU, in the caller's order.  Needs two visible devices (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)


def _two_devices():
    from multiplexnanopore_b200.engine import device_count
    return device_count() >= 2


def test_result_dict_identical_on_one_and_two_gpus():
    if not _two_devices():
        pytest.skip("needs two visible GPUs")
    from multiplexnanopore_b200 import savemoney as sm
    from multiplexnanopore_b200 import synth
    rng = np.random.default_rng(21)
    plasmids = synth.plasmid_set(rng, 4, 1500, 3200, backbone=700)
    reads, _ = synth.reads_for(rng, plasmids, 150, topology=0)
    fastq = {f"@read{i}": [r, None] for i, r in enumerate(reads)}
    one = sm.execute_alignment_core(plasmids, fastq, PARAMS, devices=[0])
    two = sm.execute_alignment_core(plasmids, fastq, PARAMS, devices=[0, 1])
    assert list(one.keys()) == list(two.keys()) == list(fastq.keys())
    for k in fastq:
        assert [r.to_dict() for r in one[k]] == [r.to_dict() for r in two[k]]
    # small batches, several in flight per GPU
    res1 = sm.align_all(plasmids, reads, PARAMS, True, devices=[0], batch_queries=20, overlap=3)
    res2 = sm.align_all(plasmids, reads, PARAMS, True, devices=[1, 0], batch_queries=20, overlap=3)
    assert np.array_equal(res1.score, res2.score) and np.array_equal(res1.status, res2.status)
    assert np.array_equal(res1.cigar_off, res2.cigar_off)
    assert all(np.array_equal(res1.runs(i), res2.runs(i)) for i in range(len(res1.score)))


def test_distance_matrix_identical_on_one_and_two_gpus():
    if not _two_devices():
        pytest.skip("needs two visible GPUs")
    from multiplexnanopore_b200 import savemoney as sm
    from multiplexnanopore_b200 import synth
    maps = synth.config_c3(seed=5, n_maps=10, families=3, len_lo=1200, len_hi=2600)
    d1 = sm.set_distance_matrix(maps, PARAMS, devices=[0])
    d2 = sm.set_distance_matrix(maps, PARAMS, devices=[0, 1])
    assert np.array_equal(d1, d2) and (d1 == d1.T).all()
