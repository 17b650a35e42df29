"""bench.py's reference arm and workload builders on the CPU (small sizes): the JSON line carries the contract's keys for
every config, the workloads are seeded (same draw twice), strong scaling deals ONE workload, weak scaling gives every rank
its own draw."""
import json
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _args(**kw):
    base = dict(config="c2", scaling="weak", reads=0, maps=0, mixtures=0, batch=1000, overlap=6)
    base.update(kw)
    return SimpleNamespace(**base)


@pytest.mark.parametrize("cfg,extra", [("c2", ["--reads", "8"]), ("c3", ["--maps", "16"]), ("c4", ["--reads", "6"]), ("legacy_dense", ["--reads", "2"])])
def test_reference_arm_line(cfg, extra):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", cfg, "--steps", "1", "--warmup", "0", "--cpu-reads", "4"] + extra,
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["higher_is_better"] is True
    assert line["gpu_launches"] == 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == line["value"]
    assert line["config"]["config"] == cfg and "workload" in line["config"]
    if cfg != "c3":
        assert line["value"] > 0


def test_workloads_are_seeded_and_sharded():
    sys.path.insert(0, str(ROOT))
    import bench
    a = bench.make_workload(_args(reads=12), 0, 1)
    b = bench.make_workload(_args(reads=12), 0, 1)
    assert a.reads == b.reads and a.refs == b.refs and len(a.reads) == 12 and len(a.refs) == 6
    other = bench.make_workload(_args(reads=12), 3, 8)                      # weak: rank 3 draws its own reads from the same mixture
    assert other.refs == a.refs and other.reads != a.reads
    strong = bench.make_workload(_args(reads=12, scaling="strong"), 3, 8)   # strong: every rank sees the ONE workload
    assert strong.reads == a.reads
    c5 = [bench.make_workload(_args(config="c5", scaling="strong", mixtures=4, reads=5), r, 2) for r in range(2)]
    assert [len(w.mixtures) for w in c5] == [2, 2] and c5[0].mixtures[0][0] != c5[1].mixtures[0][0]
    c1 = bench.make_workload(_args(config="c1", reads=3), 0, 1)
    assert sorted(len(r) for r in c1.refs) == [4722, 4877, 4916, 5107, 9884, 9933]   # the six bundled demo maps
    c3 = bench.make_workload(_args(config="c3", maps=16), 0, 1)
    assert c3.kind == "pairs" and len(c3.refs) == 16
    tasks, what = bench.cpu_tasks(c3, 10)
    assert len(tasks) == 10 and "pairs" in what
