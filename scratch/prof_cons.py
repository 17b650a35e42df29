"""one call of the Bayesian consensus on a 3 000-read x 8 000-column stack (the shape of tests/test_consensus_gpu.py's
full-depth test), for ncu: kernel time by CUDA-free wall clock of the call and the ncu capture of smx_consensus_kernel"""
import gzip, json, sys, time
sys.path.insert(0, ".")
import numpy as np
from multiplexnanopore_b200 import Engine, consensus as cs
from multiplexnanopore_b200.msa import MyMSA
t = json.load(gzip.open("tests/golden/consensus_v1.json.gz", "rt"))["tables"][0]
pdf = cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"])
rng = np.random.default_rng(41)
R, C = 3000, 8000
ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=C)
seq = np.repeat(ref[None, :], R, axis=0)
u = rng.random((R, C))
sub = u < 0.04
seq[sub] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(sub.sum()))
dele = (u >= 0.04) & (u < 0.07)
seq[dele] = ord("-")
qs = rng.integers(1, 42, size=(R, C)).astype(np.int8); qs[dele] = -1
cig = np.full((R, C), ord("="), np.uint8); cig[dele] = ord("D")
m = MyMSA(ref, seq, qs, cig, dict(error_rate=1e-7, ins_rate=1e-7))
with Engine(0) as eng:
    for k in range(2):
        t0 = time.perf_counter()
        cs.consensus_error_rates(m, pdf, engine=eng)
        print(f"call {k}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
