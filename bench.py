"""Benchmark of the alignment stage of savemoney.post_analysis on BASELINE.json configs[1] (C2):
6 synthetic plasmids x 5-10 kb, 20 000 ONT-like reads (~8 % error), circular topology.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One step = one pass of the hot path (scan -> select -> chain -> DP + traceback -> stitch for every
read x plasmid x strand) over the whole read set of the rank.  Metric = GCUPS: giga DP-cell updates
per second, cells counted as the reference algorithm counts them (sum of len(s1) * len(s2) over every
parasail call it would make) -- both arms count the same cells for the same reads.

  value        device-resident throughput: cells / (CUDA-event time of the path's kernels), measured in one extra
               pass with the batches serialised (with several batches in flight the event intervals of one handle
               contain the other handles' kernels); serialised launches cannot fill each other's tails, so value
               can come out BELOW e2e
  e2e.value    through the C-ABI call with HOST buffers (six 1000-read batches in flight), host<->device copies
               and host stages inside
  ms_per_step  wall clock per step (the e2e path), max over ranks
Multi-GPU: one process per GPU, every rank aligns its own 20k-read mixture (reads shard with no
data-path collective) => weak scaling.
"""
from __future__ import annotations

import argparse
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

PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
# Algorithmic integer lane-instructions per DP cell WITH direction bits:
#   int32 lanes: 6 score ops + 8 trace ops = 14 per cell
#   packed u16x2 lanes, diagonal frame (smx_dp16h.cuh): per cell PAIR 4 max + 1 subst + 2 add + 8 flag updates = 15
OPS_PER_CELL_INT32 = 14.0
OPS_PER_CELL_PACKED = 7.5
DPX_PER_CELL_PACKED = 2.0  # VIMNMX.U16x2 per cell (4 per cell pair)
TRACE_BYTES_PER_CELL = 0.5


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


def workload(rank: int, n_reads: int):
    """C2: the 6-plasmid mixture of seed 2; rank 0 aligns the canonical 20 000 reads, every other rank its own
    20 000 reads drawn from the SAME mixture (read-sharded weak scaling: equal work per GPU)."""
    from multiplexnanopore_b200 import synth
    plasmids, reads, truth = synth.config_c2(seed=2, n_reads=n_reads if rank == 0 else 0)
    if rank != 0:
        reads, truth = synth.reads_for(np.random.default_rng(2 + 1000 * rank), plasmids, n_reads, topology=0)
    return plasmids, reads, truth


# ------------------------------------------------------------------------------------------------
# CPU side: the reference algorithm (oracle port) on the box's host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    plasmids, read = args
    from oracle import savemoney_port as port
    log = port.DPLog()
    t = time.time()
    port.align_read(plasmids, read, PARAMS, True, log)
    return log.cells, time.time() - t


def cpu_sample(plasmids, reads, n_sample: int, cores: int):
    """aligns `n_sample` reads of the workload with the CPU oracle over `cores` processes."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    sample = reads[:n_sample]
    t0 = time.time()
    with mp.get_context("fork").Pool(processes=cores) as pool:
        out = pool.map(_cpu_worker, [(plasmids, r) for r in sample], chunksize=1)
    wall = time.time() - t0
    cells = sum(c for c, _ in out)
    return {"value": cells / wall / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
            "sample": f"first {len(sample)} of the {len(reads)} reads of the workload, all 6 plasmids x 2 strands, "
                      f"oracle/savemoney_port.py + oracle/parasail_port.c over a {cores}-process pool (the reference's "
                      f"multiprocessing.Pool(n_cpu) pattern); real parasail is not installable offline",
            "reads_per_s": len(sample) / wall, "wall_s": wall, "dp_cells": cells}


def run_reference(args, rank, world):
    if rank != 0:
        return
    plasmids, reads, _ = workload(0, args.reads)
    cores = usable_cores()
    n_sample = args.cpu_reads or max(4 * cores, 16)   # four reads per worker process: ~10 s of CPU work per step
    for _ in range(args.warmup):
        cpu_sample(plasmids, reads, min(2, n_sample), min(2, cores))
    vals, walls = [], []
    last = None
    for _ in range(args.steps):
        last = cpu_sample(plasmids, reads, n_sample, cores)
        vals.append(last["dp_cells"]); walls.append(last["wall_s"])
    v = sum(vals) / sum(walls) / 1e9
    last["value"] = v
    line = {"impl": "reference", "metric": "alignment-stage DP throughput (post_analysis, C2)", "value": v, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(args, sample_reads=n_sample), "cpu_baseline": last,
            "reads_per_s": n_sample * args.steps / sum(walls),
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(args, **extra):
    d = {"workload": "BASELINE.json configs[1]: synthetic 6 plasmids x 5-10 kb (shared 2.5 kb backbone), "
                     f"{args.reads} ONT-like reads (~8% error), circular, post_analysis alignment stage",
         "reads_per_rank": args.reads, "plasmids": 6, "pair_strands_per_rank": args.reads * 12,
         "scoring": [3, 1, 1, -2], "cache": "inputs_exceed_l2 (per-step direction-bit arena is GBs >> 126 MB L2)",
         "batch_reads_per_call": args.batch, "batches_in_flight": args.overlap}
    d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def run_engine(args, rank, world, local_rank):
    import torch
    from multiplexnanopore_b200 import Engine
    from multiplexnanopore_b200 import savemoney as sm

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    eng = Engine(local_rank)
    plasmids, reads, _ = workload(rank, args.reads)
    cores = usable_cores()
    # host stages need about three cores per GPU (measured); more threads per handle only add scheduling noise
    n_threads = args.threads or max(1, min(4, (cores // max(world, 1)) // 2))
    refs = [p.upper() for p in plasmids]
    reads = [r.upper() for r in reads]
    ref_bytes, read_bytes = sum(len(p) for p in refs), sum(len(r) for r in reads)

    def step(overlap=None):
        return sm.align_all(refs, reads, PARAMS, True, engine=eng, n_threads=n_threads, batch_queries=args.batch,
                            overlap=args.overlap if overlap is None else overlap)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    agg = None
    out_bytes = 0
    for _ in range(args.steps):
        res = step()
        out_bytes = res.nbytes
        agg = dict(res.stats) if agg is None else {k: agg[k] + res.stats[k] for k in agg}
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    if dist is not None:
        dist.barrier()
    cells = agg["dp_cells"]
    timed_stats = agg
    # host -> device per step (rank 0): every read once and the plasmid maps once per batch (the pool -- both strands,
    # each stored twice so that rotated slices are contiguous -- is built on the device), the 24-byte expansion items,
    # the DP descriptors (48 B + two 4 B orders per problem)
    n_batches = -(-len(reads) // args.batch)
    h2d_bytes = read_bytes + ref_bytes * n_batches + 24 * (2 * len(reads) + len(refs) * n_batches) + 56 * agg["n_dp_problems"] / args.steps
    if args.overlap > 1:
        # Device-time pass: with several batches in flight the CUDA-event intervals of one handle also contain
        # the other handles' kernels, so the per-kernel times behind `value` and `roofline` come from one extra
        # pass over the same inputs with the batches serialised (not part of the K timed steps; one direction-bit
        # chunk per batch).  Serialised kernels cannot fill each other's tails, so `value` can be BELOW e2e.
        eng.set_trace_budget(40 << 30)
        one = step(overlap=1)
        eng.set_trace_budget(16 << 30)
        agg = {k: v * args.steps for k, v in one.stats.items()}
    dev_s = (agg["gpu_scan_ms"] + agg["gpu_select_ms"] + agg["gpu_dp_ms"]) * 1e-3
    if dist is not None:
        t = torch.tensor([wall, dev_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_s = float(t[0]), float(t[1])
        c = torch.tensor([cells, agg["scan_cells"], float(len(reads) * args.steps), timed_stats["kernel_launches"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_all, scan_all, reads_all, launches_all = (float(x) for x in c)
    else:
        cells_all, scan_all, reads_all, launches_all = cells, agg["scan_cells"], float(len(reads) * args.steps), timed_stats["kernel_launches"]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # dominant kernel = the forward DP class with the largest device time.  Classes: 6 = int32 warp-per-problem K=8,
    # 8 = int32 CTA-per-problem, 10..13 = packed s16x2 with 1/2/4/8 warps per problem (two cells per DPX lane-instruction).
    names = {6: "smx_dp_forward<K=8,int32,warp-per-problem>", 8: "smx_dp_forward<K=8,int32,CTA-per-problem>",
             10: "smx_dp_forward16h<W=1> (u16x2 diagonal frame, warp per problem)", 11: "smx_dp_forward16h<W=2> (u16x2 diagonal frame, 2 warps per problem)",
             12: "smx_dp_forward16h<W=4>", 13: "smx_dp_forward16h<W=8>"}
    dom = max(names, key=lambda c: agg[f"dp_ms_class{c}"])
    k8_ms = agg[f"dp_ms_class{dom}"]
    k8_cells = agg[f"dp_cells_class{dom}"]
    peak = eng.measure_int_peak(3000)
    # both integer pipes together (IADD3/LOP3/VIMNMX on the ALU pipe, IMAD on the FMA pipe): the issue-rate ceiling
    int_peak_tops = peak["iadd_lop_gops"] / 1e3
    dpx_peak_tops = max(peak["viaddmax_gops"], peak["vimax3_gops"]) / 1e3
    ops_per_cell = OPS_PER_CELL_PACKED if dom >= 10 else OPS_PER_CELL_INT32
    k8_tcups = k8_cells / (k8_ms * 1e-3) / 1e12 if k8_ms > 0 else 0.0
    hbm_peak, hbm_src = measured_peaks()
    n_launch_k8 = max(1, int(agg[f"dp_launches_{dom}"]))
    roofline = {
        "bound": "alu",  # integer issue rate (ALU + FMA pipes) binds this kernel, not HBM and not tensor cores
        "kernel": names[dom] + " -- forward affine-gap DP with packed direction bits",
        "achieved": k8_tcups * ops_per_cell, "peak": int_peak_tops, "unit": "Tops/s (integer lane-instructions)",
        "frac": k8_tcups * ops_per_cell / int_peak_tops if int_peak_tops else None,
        "peak_source": "measured live by smx_measure_int_peak: independent IADD3+LOP3 chains on all SMs (ALU and FMA pipes together, 128 lanes/clk/SM)",
        "algorithmic_ops_per_cell": ops_per_cell, "cells_per_launch": k8_cells / n_launch_k8, "launches": n_launch_k8,
        "kernel_ms_per_launch": k8_ms / n_launch_k8, "kernel_tcups": k8_tcups,
        "note": "cells are the reference's m x n per problem; the kernel also computes the band padding (17 % of the cells on this workload before lone last bands were moved to 4 rows per lane); the launch time is measured with the batches serialised, where it is bound by the largest problems of the batch (one warp each), so this is a lower bound of the sustained rate (DESIGN.md 3.1)",
        "dpx_view": {"achieved_tops": k8_tcups * DPX_PER_CELL_PACKED, "peak_tops": dpx_peak_tops,
                     "frac": k8_tcups * DPX_PER_CELL_PACKED / dpx_peak_tops if dpx_peak_tops else None,
                     "note": "VIMNMX.U16x2 lane-instructions against the measured DPX rate of the ALU pipe (64 lanes/clk/SM)"},
        "hbm_view": {"achieved": k8_tcups * 1e3 * TRACE_BYTES_PER_CELL, "peak": hbm_peak, "unit": "GB/s", "frac": k8_tcups * 1e3 * TRACE_BYTES_PER_CELL / hbm_peak,
                     "peak_source": f"MEASURED_PEAKS.json ({hbm_src})"},
        "traffic": None,
        "int_peak_detail_gops": peak,
    }
    traffic_file = ROOT / "profiles" / "dp_forward_traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text())
        except Exception:
            pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:  # rank 0 at N = 1 only
        cpu = cpu_sample(plasmids, reads, args.cpu_reads or max(4 * cores, 16), cores)
    line = {
        "metric": "alignment-stage DP throughput (post_analysis, C2)", "value": cells_all / dev_s / 1e9, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16x2 (int32 for problems of <= 32 rows)", "data": "synthetic",
        "config": config_dict(args, host_threads_per_rank=n_threads, host_cores=cores),
        "e2e": {"value": cells_all / wall / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(out_bytes),
                "reads_per_s": reads_all / wall,
                "note": "smx_pipeline_run from host buffers: pool upload, descriptors, region lists, CIGAR runs and results cross PCIe inside the timed region"},
        "reads_per_s_device": reads_all / dev_s, "scan_gcups_device": scan_all / (agg["gpu_scan_ms"] * 1e-3) / 1e9 if agg["gpu_scan_ms"] else None,
        "device_ms_per_step": 1e3 * dev_s / args.steps,
        "device_timing": "CUDA events on the engine stream, one extra serialised pass" if args.overlap > 1 else "CUDA events on the engine stream, timed steps",
        "stage_seconds_per_step_rank0": {k: agg[k] / args.steps for k in ("t_pool", "t_scan", "t_select", "t_chain", "t_dp", "t_stitch", "t_total")},
        "stage_seconds_per_step_rank0_overlapped": {k: timed_stats[k] / args.steps for k in ("t_pool", "t_scan", "t_select", "t_chain", "t_dp", "t_stitch", "t_total")},
        "kernel_ms_per_step_rank0": {"scan": agg["gpu_scan_ms"] / args.steps, "select_and_chain": agg["gpu_select_ms"] / args.steps, "select": agg["gpu_select_only_ms"] / args.steps, "chain": agg["gpu_chain_ms"] / args.steps,
                                     "dp_all": agg["gpu_dp_ms"] / args.steps, "dp_forward16_w8": agg["dp_ms_class13"] / args.steps, "dp_forward16_w4": agg["dp_ms_class12"] / args.steps,
                                     "dp_forward16_w2": agg["dp_ms_class11"] / args.steps, "dp_forward16_w1": agg["dp_ms_class10"] / args.steps,
                                     "dp_forward_k8_team": agg["dp_ms_class8"] / args.steps,
                                     "dp_forward_k8_solo": agg["dp_ms_class6"] / args.steps,
                                     "dp_forward_k4": agg["dp_ms_class4"] / args.steps, "dp_forward_k2": agg["dp_ms_class2"] / args.steps,
                                     "dp_forward_k1": agg["dp_ms_class0"] / args.steps, "traceback": agg["dp_ms_traceback"] / args.steps,
                                     "compact": agg["dp_ms_compact"] / args.steps},
        "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "parity": "bit-exact vs reference-generated goldens (tests/golden) and the CPU oracle; parasail tie rules unpinned (DESIGN.md)",
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--reads", type=int, default=20000, help="reads per rank (C2 = 20000)")
    ap.add_argument("--batch", type=int, default=1000, help="reads per smx_pipeline_run call")
    ap.add_argument("--overlap", type=int, default=6, help="batches in flight per rank (engine handles)")
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads of the CPU sample (default: max(4 * cores, 16))")
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
