"""Benchmark of the alignment stage.  Default = BASELINE.json configs[1] (C2): post_analysis on
6 synthetic plasmids x 5-10 kb, 20 000 ONT-like reads (~8 % error), circular topology.

    python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA engine, C2, weak scaling
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on the box's host cores
    python bench.py --config c1|c2|c3|c4|c5|legacy_dense            # the other BASELINE.json configs (own runs, profiles/)
    python bench.py --scaling strong                                # ONE workload sharded over the N ranks

One step = one pass of the hot path (scan -> select -> chain -> DP + traceback -> stitch for every read x plasmid x
strand) over the whole unit set of the rank.  Metric = GCUPS: giga DP-cell updates per second, cells counted as the
reference algorithm counts them (sum of len(s1) * len(s2) over every parasail call it would make, SURVEY.md section
8d) -- both arms count the same cells for the same inputs.

  value        device-resident throughput: cells / (CUDA-event time of the path's kernels: scan, select, chain, forward
               DP, traceback, CIGAR compaction), measured in one extra pass with the batches serialised (with several
               batches in flight the event intervals of one handle contain the other handles' kernels).  Serialised
               launches cannot fill each other's tails, so value can come out BELOW e2e.  `value_source` says so.
  e2e.value    through the C-ABI call with HOST buffers (eight 1000-read batches in flight), host<->device copies and
               host stages inside the timed region: the headline
  ms_per_step  wall clock per step (the e2e path), max over ranks
Multi-GPU: one process per GPU.  weak: every rank aligns its own draw of the workload; strong: the units (reads, plasmid
pairs, mixtures) of ONE workload are dealt to the ranks (shard.deal), the score-matrix rows are gathered to rank 0
through NCCL inside the timed region, the CIGARs stay with the rank that feeds them to the consensus stage.
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before torch / the engine create the CUDA context (multiplexnanopore_b200/_lib.py)

PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
# Algorithmic integer lane-instructions per DP cell WITH direction bits:
#   int32 lanes: 6 score ops + 8 trace ops = 14 per cell
#   packed u16x2 lanes, diagonal frame (smx_dp16h.cuh): per cell PAIR 4 max + 1 subst + 2 add + 8 flag updates = 15
OPS_PER_CELL_INT32 = 14.0
OPS_PER_CELL_PACKED = 7.5
DPX_PER_CELL_PACKED = 2.0  # VIMNMX.U16x2 per cell (4 per cell pair)
# checkpointed problems (SMX_MODE_CKPT): the forward launch computes scores only -- per cell PAIR 2 max + 1 max3 + 1 subst + 2 add = 6
OPS_PER_CELL_SCORE_ONLY = 3.0
DPX_PER_CELL_SCORE_ONLY = 1.5
CK_BYTES_PER_CELL = 0.19   # 22 words x 32 lanes per 32-step chunk and 512-row band pair + one bottom-line entry per 256 cells
TRACE_BYTES_PER_CELL = 0.5
CPU_SAMPLE_READS = 200     # seeded 200-read subset for the cpu_baseline of the engine arm


def usable_cores() -> int:
    """cores this process may run on (taskset / cgroup cpusets), not the machine's core count"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        util = [float(r[7]) for r in self.rows if len(r) > 7 and r[7].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "gpu_util_pct_mean": float(np.mean(util)) if util else None,
                "power_w_mean": float(np.mean(pw)) if pw else None}


# ------------------------------------------------------------------------------------------------
# Workloads: BASELINE.json configs (synthetic, seeded)
# ------------------------------------------------------------------------------------------------
class Workload:
    """kind 'reads': refs + reads (post_analysis);  'pairs': maps, all i<j pairs (pre_survey);
    'mixtures': list of (refs, reads) (batch post_analysis);  'dense': refs + reads, legacy sw_trace(read, ref+ref)"""

    def __init__(self, name, kind, description, params, **kw):
        self.name, self.kind, self.description, self.params = name, kind, description, params
        self.__dict__.update(kw)


def demo_maps():
    """the six bundled demo plasmid maps (resources/demo_data/my_plasmid_maps_fa/*.fa), carried by the golden fixture
    because /root/reference does not travel to the GPU box"""
    with gzip.open(ROOT / "tests" / "golden" / "golden_v1.json.gz", "rt") as f:
        return list(json.load(f)["demo_maps"].values())


def make_workload(args, rank: int, world: int) -> Workload:
    from multiplexnanopore_b200 import synth
    cfg, strong = args.config, args.scaling == "strong"
    # weak scaling: rank 0 the canonical draw, every other rank its own draw of the same size from the same generator
    draw = 0 if (strong or rank == 0) else rank
    if cfg == "c1":
        refs = demo_maps()
        n = args.reads or 2000
        reads, _ = synth.reads_for(np.random.default_rng(1 + 1000 * draw), refs, n, topology=0)
        return Workload("c1", "reads", "BASELINE.json configs[0]: the six bundled demo plasmid maps (4.7-9.9 kb) + "
                        f"{n} synthetic ONT-like reads (the demo FASTQ is a missing blob), circular, post_analysis alignment stage",
                        dict(PARAMS), refs=refs, reads=reads)
    if cfg == "c2":
        n = args.reads or 20000
        plasmids, reads, _ = synth.config_c2(seed=2, n_reads=n if draw == 0 else 0)
        if draw != 0:
            reads, _ = synth.reads_for(np.random.default_rng(2 + 1000 * draw), plasmids, n, topology=0)
        return Workload("c2", "reads", "BASELINE.json configs[1]: synthetic 6 plasmids x 5-10 kb (shared 2.5 kb backbone), "
                        f"{n} ONT-like reads (~8% error), circular, post_analysis alignment stage", dict(PARAMS), refs=plasmids, reads=reads)
    if cfg == "c3":
        n = args.maps or 96
        maps = synth.config_c3(seed=3 + 1000 * draw, n_maps=n)
        return Workload("c3", "pairs", f"BASELINE.json configs[2]: pre_survey all-pairs distance matrix for {n} synthetic plasmid maps "
                        f"(8 families, 2-15 kb), circular: {n * (n - 1) // 2} pairs x 2 strands", dict(PARAMS), refs=maps)
    if cfg == "c4":
        n = args.reads or 20000
        rng = np.random.default_rng(4)
        plasmids = synth.plasmid_set(rng, 12, 3000, 20000, backbone=2000)
        reads, _ = synth.reads_for(np.random.default_rng(4 + 1000 * draw + 7), plasmids, n, topology=1)
        return Workload("c4", "reads", "BASELINE.json configs[3]: synthetic 12-plasmid mixture (3-20 kb), linear topology, score matrix + traceback; "
                        f"{n} of the 200 000 reads per step (same generator, bounded so that a step takes seconds)",
                        dict(PARAMS, topology_of_dna=1), refs=plasmids, reads=reads)
    if cfg == "c5":
        m = args.mixtures or (96 if world == 8 else 12 * world)
        n = args.reads or 20000
        mixtures = []
        for k in range(m):
            if strong and k % world != rank:   # mixtures are dealt to the ranks (batch mode = independent post_analysis runs)
                continue
            seed = 100 + (k if strong else k + 1000 * draw)
            rng = np.random.default_rng(seed)
            pl = synth.plasmid_set(rng)
            rd, _ = synth.reads_for(rng, pl, n, topology=0)
            mixtures.append((pl, rd))
        return Workload("c5", "mixtures", f"BASELINE.json configs[4]: batch post_analysis, {m} synthetic mixtures x 6 plasmids x {n} reads "
                        "(seeds 100..), mixtures dealt to the GPUs", dict(PARAMS), mixtures=mixtures, n_mixtures_total=m)
    if cfg == "legacy_dense":
        n = args.reads or 256
        plasmids, reads, _ = synth.config_c2(seed=2, n_reads=n if draw == 0 else 0)
        if draw != 0:
            reads, _ = synth.reads_for(np.random.default_rng(2 + 1000 * draw), plasmids, n, topology=0)
        return Workload("legacy_dense", "dense", "legacy dense mode (Modules/1_alignment_consensus_core.py:219-255): sw_trace(read, ref + ref) per "
                        f"read x plasmid x strand, C2 plasmids, {n} reads", dict(PARAMS), refs=plasmids, reads=reads)
    raise SystemExit(f"unknown --config {cfg}")


# ------------------------------------------------------------------------------------------------
# CPU side: the reference algorithm (oracle port) on the box's host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(task):
    kind, params, a, b = task
    from oracle import savemoney_port as port
    log = port.DPLog()
    if kind == "read":
        port.align_read(a, b, params, True, log)
    elif kind == "pair":
        port.distance(a, b, params, log)
    else:  # dense: sw_trace(read, ref + ref) on both strands of every plasmid (the legacy MyAligner.align_all)
        from oracle import oracle as orc
        cells = 0
        for ref in a:
            for q in (b, port.revcomp(b)):
                orc.align(orc.MODE_SW, q, ref + ref, params["gap_open_penalty"], params["gap_extend_penalty"], params["match_score"], params["mismatch_score"])
                cells += len(q) * 2 * len(ref)
        return cells
    return log.cells


def cpu_tasks(wl: Workload, n_sample: int):
    if wl.kind == "reads":
        return [("read", wl.params, wl.refs, r) for r in wl.reads[:n_sample]], f"first {min(n_sample, len(wl.reads))} of the {len(wl.reads)} reads of the workload, all {len(wl.refs)} plasmids x 2 strands"
    if wl.kind == "pairs":
        n = len(wl.refs)
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
        step = max(1, len(pairs) // n_sample)   # evenly spaced pairs: same-family and cross-family in proportion
        sel = pairs[::step][:n_sample]
        return [("pair", wl.params, wl.refs[i], wl.refs[j]) for i, j in sel], f"{len(sel)} evenly spaced pairs of the {len(pairs)} plasmid pairs, both strands"
    if wl.kind == "mixtures":
        pl, rd = wl.mixtures[0]
        return [("read", wl.params, pl, r) for r in rd[:n_sample]], f"first {min(n_sample, len(rd))} reads of the first mixture of the rank, 6 plasmids x 2 strands"
    return [("dense", wl.params, wl.refs, r) for r in wl.reads[:n_sample]], f"first {min(n_sample, len(wl.reads))} reads, sw_trace(read, ref + ref) x {len(wl.refs)} plasmids x 2 strands"


def cpu_sample(wl: Workload, n_sample: int, cores: int):
    """runs a bounded sample of the workload with the CPU oracle over `cores` processes."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    tasks, what = cpu_tasks(wl, n_sample)
    t0 = time.time()
    with mp.get_context("fork").Pool(processes=cores) as pool:
        out = pool.map(_cpu_worker, tasks, chunksize=1)
    wall = time.time() - t0
    cells = sum(out)
    return {"value": cells / wall / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
            "sample": f"{what}; oracle/savemoney_port.py + oracle/parasail_port.c over a {cores}-process pool (the reference's "
                      f"multiprocessing.Pool(n_cpu) pattern); real parasail is not installable offline (profiles/parasail_probe_r02.log)",
            "units_per_s": len(tasks) / wall, "reads_per_s": len(tasks) / wall if wl.kind != "pairs" else None, "wall_s": wall, "dp_cells": cells,
            "n_units": len(tasks)}


def config_dict(args, wl: Workload, **extra):
    d = {"workload": wl.description, "config": wl.name, "scoring": [3, 1, 1, -2],
         "cache": "inputs_exceed_l2 (per-step direction-bit arena is GBs >> 126 MB L2)",
         "batch_reads_per_call": args.batch, "batches_in_flight": args.overlap}
    if wl.kind in ("reads", "dense"):
        d.update(reads_per_rank=len(wl.reads), plasmids=len(wl.refs), pair_strands_per_rank=len(wl.reads) * 2 * len(wl.refs))
    elif wl.kind == "pairs":
        n = len(wl.refs)
        d.update(maps=n, pair_strands=n * (n - 1))
    else:
        d.update(mixtures_total=wl.n_mixtures_total, mixtures_this_rank=len(wl.mixtures))
    d.update(extra)
    return d


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = make_workload(args, 0, 1 if args.scaling == "weak" else world)
    cores = usable_cores()
    n_sample = args.cpu_reads or max(4 * cores, 16)   # four units per worker process: ~5-10 s of CPU work per step
    for _ in range(args.warmup):
        cpu_sample(wl, min(2, n_sample), min(2, cores))
    cells, walls, last = [], [], None
    for _ in range(args.steps):
        last = cpu_sample(wl, n_sample, cores)
        cells.append(last["dp_cells"]); walls.append(last["wall_s"])
    v = sum(cells) / sum(walls) / 1e9
    last["value"] = v
    line = {"impl": "reference", "metric": f"alignment-stage DP throughput ({wl.name})", "value": v, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(args, wl, sample_units=last["n_units"]), "cpu_baseline": last,
            "units_per_s": last["n_units"] * args.steps / sum(walls),
            "reads_per_s": last["n_units"] * args.steps / sum(walls) if wl.kind != "pairs" else None,
            "note": "each step is a bounded sample of the workload: value is a RATE (cells per second), comparable with the engine arm's rate",
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def run_engine(args, rank, world, local_rank):
    import torch
    from multiplexnanopore_b200 import Engine, shard
    from multiplexnanopore_b200 import savemoney as sm

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    eng = Engine(local_rank)
    wl = make_workload(args, rank, world)
    strong = args.scaling == "strong" and world > 1
    cores = usable_cores()
    # host stages need about three cores per GPU (measured); more threads per handle only add scheduling noise
    n_threads = args.threads or max(1, min(4, (cores // max(world, 1)) // 2))
    params = wl.params
    up = lambda seqs: [s.upper() for s in seqs]

    # ---- the units of this rank ----
    n_units_total = 0
    if wl.kind in ("reads", "dense"):
        refs, reads = up(wl.refs), up(wl.reads)
        n_units_total = len(reads)
        mine = shard.shard_reads([len(r) for r in reads], world, rank) if strong else np.arange(len(reads))
        my_reads = [reads[i] for i in mine]
        in_bytes = sum(len(p) for p in refs) * (-(-len(my_reads) // args.batch)) + sum(len(r) for r in my_reads)
    elif wl.kind == "pairs":
        refs = up(wl.refs)
        n = len(refs)
        all_pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], dtype=np.int32).reshape(-1, 2)
        n_units_total = len(all_pairs)
        mine = shard.deal([len(refs[i]) * len(refs[j]) for i, j in all_pairs], world)[rank] if strong else np.arange(len(all_pairs))
        my_pairs = all_pairs[mine]
        in_bytes = sum(len(p) for p in refs)
    else:
        mixtures = [(up(pl), up(rd)) for pl, rd in wl.mixtures]
        n_units_total = wl.n_mixtures_total if strong else len(mixtures)
        in_bytes = sum(sum(len(p) for p in pl) * (-(-len(rd) // args.batch)) + sum(len(r) for r in rd) for pl, rd in mixtures)

    def step(overlap=None):
        ov = args.overlap if overlap is None else overlap
        if wl.kind == "reads":
            return [sm.align_all(refs, my_reads, params, True, engine=eng, n_threads=n_threads, batch_queries=args.batch, overlap=ov)]
        if wl.kind == "pairs":
            return [sm.align_all(refs, refs, params, False, pairs=my_pairs, engine=eng, n_threads=max(n_threads, min(8, cores // max(world, 1))))]
        if wl.kind == "mixtures":
            return [sm.align_all(pl, rd, params, True, engine=eng, n_threads=n_threads, batch_queries=args.batch, overlap=ov) for pl, rd in mixtures]
        return dense_step(2 if overlap is None else overlap)

    dense_state = {}

    def dense_step(calls_in_flight):
        fastq = {i: [r, None] for i, r in enumerate(my_reads)}
        t0 = time.perf_counter()
        out = sm.legacy_dense_align(refs, fastq, params, engine=eng, stats=dense_state, overlap=calls_in_flight)
        dense_state["wall"] = time.perf_counter() - t0
        return [out]

    def gather_scores(results):
        """strong scaling: the rows of the score matrix travel to rank 0 (the only inter-GPU step)"""
        if not strong or wl.kind not in ("reads", "pairs"):
            return None
        sc = torch.from_numpy(np.ascontiguousarray(results[0].score)).cuda()
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([sc.numel()], dtype=torch.int64, device="cuda"))
        cap = int(max(int(s) for s in sizes))
        buf = torch.zeros(cap, dtype=torch.int32, device="cuda"); buf[: sc.numel()] = sc
        out = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None
        dist.gather(buf, out, dst=0)
        if rank == 0:
            return [o[: int(s)].cpu().numpy() for o, s in zip(out, sizes)]
        return None

    for _ in range(args.warmup):
        gather_scores(step())
    sampler = ClockSampler(local_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    agg, out_bytes = None, 0
    for _ in range(args.steps):
        results = step()
        gather_scores(results)
        if wl.kind == "dense":
            st = {"dp_cells": dense_state["cells"], "kernel_launches": dense_state["launches"], "n_dp_problems": dense_state["problems"]}
            out_bytes = dense_state["out_bytes"]
        else:
            out_bytes = sum(r.nbytes for r in results)
            st = {}
            for r in results:
                for k, v in r.stats.items():
                    st[k] = st.get(k, 0) + v
        agg = dict(st) if agg is None else {k: agg[k] + st[k] for k in agg}
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    if dist is not None:
        dist.barrier()
    cells = agg["dp_cells"]
    timed_stats = agg
    h2d_bytes = in_bytes + 56 * agg.get("n_dp_problems", 0) / args.steps

    # ---- device-time pass ----
    value_source = "CUDA events on the engine streams, timed steps"
    if wl.kind == "dense":
        step(overlap=1)   # one extra pass with the calls serialised: CUDA-event intervals of one handle then hold only its own kernels
        dev_s = dense_state["kernel_ms_total"] * 1e-3 * args.steps
        agg = dict(agg, **{f"dp_ms_class{c}": dense_state["class_ms"][c] * args.steps for c in range(14)},
                   **{f"dp_cells_class{c}": dense_state["class_cells"][c] * args.steps for c in range(14)},
                   **{f"dp_launches_{c}": dense_state["class_launches"][c] * args.steps for c in range(16)},
                   dp_ms_traceback=dense_state["class_ms"][14] * args.steps, dp_ms_compact=dense_state["class_ms"][15] * args.steps)
        value_source = "one extra pass with the calls serialised, outside the K timed steps: CUDA events around every forward / traceback / compaction launch, scaled to the K steps"
    else:
        if args.overlap > 1 and wl.kind != "pairs":
            # with several batches in flight the CUDA-event intervals of one handle also contain the other handles'
            # kernels, so the per-kernel times behind `value` and `roofline` come from ONE EXTRA pass over the same
            # inputs with the batches serialised (outside the K timed steps; 40 GiB of direction bits per chunk)
            before = eng.trace_budget
            eng.set_trace_budget(40 << 30)
            one = step(overlap=1)
            eng.set_trace_budget(before)
            agg = {}
            for r in one:
                for k, v in r.stats.items():
                    agg[k] = agg.get(k, 0) + v * args.steps
            value_source = ("one extra pass with the batches serialised, outside the K timed steps (kernels only: scan + select + chain + "
                            "forward DP + traceback + CIGAR compaction; the pool expand / pack kernels, < 0.3 % of the step, are not timed)")
        dev_s = (agg["gpu_scan_ms"] + agg["gpu_select_ms"] + agg["gpu_dp_ms"] + agg["dp_ms_compact"]) * 1e-3
    units_rank = {"reads": lambda: len(my_reads), "dense": lambda: len(my_reads), "pairs": lambda: len(my_pairs),
                  "mixtures": lambda: sum(len(rd) for _, rd in mixtures)}[wl.kind]()
    if dist is not None:
        t = torch.tensor([wall, dev_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_s = float(t[0]), float(t[1])
        c = torch.tensor([cells, agg.get("scan_cells", 0), float(units_rank * args.steps), timed_stats["kernel_launches"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_all, scan_all, units_all, launches_all = (float(x) for x in c)
    else:
        cells_all, scan_all, units_all, launches_all = cells, agg.get("scan_cells", 0), float(units_rank * args.steps), timed_stats["kernel_launches"]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # dominant kernel = the forward DP class with the largest device time.  Classes: kshift * 2 + local for the int32
    # warp-per-problem kernels, 8/9 int32 CTA-per-problem global/local, 10..13 packed u16x2 with 1/2/4/8 warps per problem
    names = {0: "smx_dp_forward<K=1,int32>", 2: "smx_dp_forward<K=2,int32>", 4: "smx_dp_forward<K=4,int32>",
             6: "smx_dp_forward<K=8,int32,warp-per-problem>", 7: "smx_dp_forward<K=8,int32,local,warp-per-problem>",
             8: "smx_dp_forward<K=8,int32,CTA-per-problem>", 9: "smx_dp_forward<K=8,int32,local,CTA-per-problem>",
             10: "smx_dp_forward16h<W=1> (u16x2 diagonal frame, warp per problem)", 11: "smx_dp_forward16h<W=2> (u16x2 diagonal frame, 2 warps per problem)",
             12: "smx_dp_forward16h<W=4>", 13: "smx_dp_forward16h<W=8>"}
    dom = max(names, key=lambda c: agg.get(f"dp_ms_class{c}", 0.0))
    k8_ms = agg.get(f"dp_ms_class{dom}", 0.0)
    k8_cells = agg.get(f"dp_cells_class{dom}", 0.0)
    peak = eng.measure_int_peak(3000)
    # both integer pipes together (IADD3/LOP3/VIMNMX on the ALU pipe, IMAD on the FMA pipe): the issue-rate ceiling
    int_peak_tops = peak["iadd_lop_gops"] / 1e3
    dpx_peak_tops = max(peak["viaddmax_gops"], peak["vimax3_gops"]) / 1e3
    ops_per_cell = OPS_PER_CELL_PACKED if dom >= 10 else OPS_PER_CELL_INT32
    k8_tcups = k8_cells / (k8_ms * 1e-3) / 1e12 if k8_ms > 0 else 0.0
    # launches of the packed phase that ran score-only (checkpointed problems) count their own algorithmic instructions
    ck_cells = min(agg.get("dp_cells_ck", 0.0), k8_cells) if dom == 10 else 0.0
    ck_share = ck_cells / k8_cells if k8_cells else 0.0
    ops_blend = ops_per_cell * (1.0 - ck_share) + OPS_PER_CELL_SCORE_ONLY * ck_share
    dpx_blend = DPX_PER_CELL_PACKED * (1.0 - ck_share) + DPX_PER_CELL_SCORE_ONLY * ck_share
    tb_ms = agg.get("dp_ms_traceback", 0.0)
    dp_stage_tcups = k8_cells / ((k8_ms + tb_ms) * 1e-3) / 1e12 if k8_ms > 0 else 0.0
    hbm_peak, hbm_src = measured_peaks()
    n_launch_k8 = max(1, int(agg.get(f"dp_launches_{dom}", 1)))
    total_class_cells = sum(agg.get(f"dp_cells_class{c}", 0.0) for c in range(14)) or 1.0
    roofline = {
        "bound": "alu",  # integer issue rate (ALU + FMA pipes) binds this kernel, not HBM and not tensor cores
        "kernel": names[dom] + (" -- forward affine-gap DP: score-only launches for the checkpointed problems (3.0 lane-instructions per cell) next to the "
                                "launches that write packed direction bits (7.5 per cell), timed as one phase" if ck_share > 0 else " -- forward affine-gap DP with packed direction bits"),
        "achieved": k8_tcups * ops_blend, "peak": int_peak_tops, "unit": "Tops/s (integer lane-instructions)",
        "frac": k8_tcups * ops_blend / int_peak_tops if int_peak_tops else None,
        "peak_source": "measured live by smx_measure_int_peak: independent IADD3+LOP3 chains on all SMs (ALU and FMA pipes together, 128 lanes/clk/SM)",
        "algorithmic_ops_per_cell": ops_blend, "algorithmic_ops_per_cell_with_direction_bits": ops_per_cell, "algorithmic_ops_per_cell_score_only": OPS_PER_CELL_SCORE_ONLY,
        "checkpointed_cell_share": ck_share,
        # the whole DP stage of the step (forward phase + both traceback kernels, i.e. including the blocks smx_traceback_ck
        # computes a second time) at the instruction count of the one-pass algorithm -- the definition round 1 and the first
        # half of round 2 reported `frac` on; comparable across the change of algorithm, not a statement about one kernel
        "one_pass_equivalent": {"ops_per_cell": ops_per_cell, "ms_per_step": (k8_ms + tb_ms) / args.steps, "tcups": dp_stage_tcups,
                                "frac": dp_stage_tcups * ops_per_cell / int_peak_tops if int_peak_tops else None,
                                "forward_phase_only_frac": k8_tcups * ops_per_cell / int_peak_tops if int_peak_tops else None},
        "cells_per_launch": k8_cells / n_launch_k8, "launches": n_launch_k8,
        "kernel_ms_per_launch": k8_ms / n_launch_k8, "kernel_tcups": k8_tcups,
        "note": "cells are the reference's m x n per problem; the kernel also computes the band padding (rows up to the next 256 / 128, 95 extra steps per band pair: "
                "DESIGN.md 3.1); launch times come from the serialised pass, whose tail (the largest problems, one warp each) other batches fill in the e2e regime",
        "dpx_view": {"achieved_tops": k8_tcups * dpx_blend, "peak_tops": dpx_peak_tops,
                     "frac": k8_tcups * dpx_blend / dpx_peak_tops if dpx_peak_tops else None,
                     "note": "VIMNMX(3).U16x2 lane-instructions against the measured DPX rate of the ALU pipe (64 lanes/clk/SM); the score-only launches are bound by "
                             "that pipe (ncu: ALU pipe 78 % busy, FMA pipe 19 %, profiles/dp_forward16h_ck_r02_summary.json)"},
        "hbm_view": {"achieved": k8_tcups * 1e3 * (TRACE_BYTES_PER_CELL * (1.0 - ck_share) + CK_BYTES_PER_CELL * ck_share), "peak": hbm_peak, "unit": "GB/s",
                     "frac": k8_tcups * 1e3 * (TRACE_BYTES_PER_CELL * (1.0 - ck_share) + CK_BYTES_PER_CELL * ck_share) / hbm_peak,
                     "peak_source": f"MEASURED_PEAKS.json ({hbm_src})"},
        "traffic": None,
        "int_peak_detail_gops": peak,
        "cell_share_by_kernel_class": {names.get(c, f"class{c}"): agg.get(f"dp_cells_class{c}", 0.0) / total_class_cells for c in range(14) if agg.get(f"dp_cells_class{c}", 0.0) > 0},
    }
    traffic_file = ROOT / "profiles" / "dp_forward_traffic.json"
    if traffic_file.exists() and dom == 10:
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text())
        except Exception:
            pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:  # rank 0 at N = 1 only
        cpu = cpu_sample(wl, args.cpu_reads or CPU_SAMPLE_READS, cores)
    unit_name = {"reads": "reads", "dense": "reads", "pairs": "pairs", "mixtures": "reads"}[wl.kind]
    ms = lambda k: agg.get(k, 0.0) / args.steps
    line = {
        "metric": f"alignment-stage DP throughput ({'post_analysis' if wl.kind != 'pairs' else 'pre_survey'}, {wl.name.upper()})", "value": cells_all / dev_s / 1e9, "unit": "GCUPS",
        "value_source": value_source,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u16x2 (int32 for problems of <= 32 rows, local mode and out-of-range problems)", "data": "synthetic",
        "config": config_dict(args, wl, host_threads_per_rank=n_threads, host_cores=cores, units_total=n_units_total),
        "e2e": {"value": cells_all / wall / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(out_bytes),
                f"{unit_name}_per_s": units_all / wall, "reads_per_s": units_all / wall if unit_name == "reads" else None,
                "note": "smx_pipeline_run from host buffers: pool upload, descriptors, region lists, CIGAR runs and results cross PCIe inside the timed region"
                        + ("; strong scaling: the score rows of every rank are gathered to rank 0 over NCCL inside the timed region" if strong else "")},
        f"{unit_name}_per_s_device": units_all / dev_s, "scan_gcups_device": scan_all / (agg["gpu_scan_ms"] * 1e-3) / 1e9 if agg.get("gpu_scan_ms") else None,
        "device_ms_per_step": 1e3 * dev_s / args.steps,
        "kernel_ms_per_step_rank0": {"scan": ms("gpu_scan_ms"), "select_and_chain": ms("gpu_select_ms"), "select": ms("gpu_select_only_ms"), "chain": ms("gpu_chain_ms"),
                                     "dp_all": ms("gpu_dp_ms"), "dp_forward16_w8": ms("dp_ms_class13"), "dp_forward16_w4": ms("dp_ms_class12"),
                                     "dp_forward16_w2": ms("dp_ms_class11"), "dp_forward16_w1": ms("dp_ms_class10"),
                                     "dp_forward_k8_team": ms("dp_ms_class8"), "dp_forward_k8_team_local": ms("dp_ms_class9"),
                                     "dp_forward_k8_solo": ms("dp_ms_class6"), "dp_forward_k8_solo_local": ms("dp_ms_class7"),
                                     "dp_forward_k4": ms("dp_ms_class4"), "dp_forward_k2": ms("dp_ms_class2"),
                                     "dp_forward_k1": ms("dp_ms_class0"), "traceback": ms("dp_ms_traceback"), "compact": ms("dp_ms_compact")},
        "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "parity": "bit-exact vs reference-generated goldens (tests/golden) and the CPU oracle; parasail itself is not reachable on this pod "
                  "(profiles/parasail_probe_r02.log), its tie rules are restated (DESIGN.md section 5, profiles/tie_sensitivity_r02.json)",
    }
    if wl.kind != "dense":
        line["stage_seconds_per_step_rank0"] = {k: agg[k] / args.steps for k in ("t_pool", "t_scan", "t_select", "t_chain", "t_dp", "t_stitch", "t_total")}
        line["stage_seconds_per_step_rank0_overlapped"] = {k: timed_stats[k] / args.steps for k in ("t_pool", "t_scan", "t_select", "t_chain", "t_dp", "t_stitch", "t_total")}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "legacy_dense"], help="BASELINE.json config (default c2: the metric's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: every rank its own draw; strong: one workload sharded over the ranks")
    ap.add_argument("--reads", type=int, default=0, help="reads per workload (c2: 20000, c1: 2000, c4: 20000 of the 200000, legacy_dense: 256)")
    ap.add_argument("--maps", type=int, default=0, help="plasmid maps of c3 (96)")
    ap.add_argument("--mixtures", type=int, default=0, help="mixtures of c5 (96 at 8 GPUs, else 12 per GPU)")
    ap.add_argument("--batch", type=int, default=1000, help="reads per smx_pipeline_run call")
    ap.add_argument("--overlap", type=int, default=8, help="batches in flight per rank (engine handles)")
    ap.add_argument("--cpu-reads", type=int, default=0, help=f"units of the CPU sample (engine arm: {CPU_SAMPLE_READS}; reference arm: max(4 * cores, 16) per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads", type=int, default=0, help="host threads per rank (default: cores / ranks)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
