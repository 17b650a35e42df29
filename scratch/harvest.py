"""Harvests the DP problem list (mode, rows, cols) of C2 reads with the CPU oracle port -> scratch/harvest_c2.npy"""
import sys
from multiprocessing import Pool
import numpy as np
sys.path.insert(0, ".")
from multiplexnanopore_b200 import synth
from oracle import oracle as orc
from oracle import savemoney_port as port
PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200
orc.build()
plasmids, reads, _ = synth.config_c2(seed=2, n_reads=n_reads)
def one(r):
    log = port.DPLog(); port.align_read(plasmids, r, PARAMS, True, log); return log.calls
if __name__ == "__main__":
    with Pool(8) as p:
        calls = [c for cs in p.map(one, reads) for c in cs]
    np.save("scratch/harvest_c2.npy", np.array(calls))
    print(len(calls))
