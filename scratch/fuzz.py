"""one-off randomized parity campaign: CUDA engine vs CPU oracle port on varied inputs (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, '.')
import numpy as np
import multiprocessing as mp
from multiplexnanopore_b200 import savemoney as sm, synth, Engine
from oracle import oracle as orc, savemoney_port as port

def port_read(args):
    plasmids, read, params, omit = args
    try:
        return [(orc.rle(r.my_cigar), int(r.score), None if r.new_query_seq_offset is None else int(r.new_query_seq_offset)) for r in port.align_read(plasmids, read, params, omit)]
    except AssertionError as e:
        return "ASSERT"

def sprinkle(rng, s, letters, rate):
    a = np.frombuffer(s.encode(), dtype=np.uint8).copy()
    m = rng.random(a.size) < rate
    a[m] = np.frombuffer(letters.encode(), dtype=np.uint8)[rng.integers(0, len(letters), size=int(m.sum()))]
    return a.tobytes().decode()

def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    eng = Engine(0)
    orc.build()
    pool = mp.get_context("fork").Pool(processes=os.cpu_count())
    nbad = 0; ntot = 0; nassert = 0
    for case in range(ncase):
        rng = np.random.default_rng(seed0 + case)
        topo = int(rng.random() < 0.35)
        npl = int(rng.integers(2, 5))
        plasmids = synth.plasmid_set(rng, npl, 700, 2600, backbone=int(rng.integers(200, 700)))
        kind = case % 6
        if kind == 1:   # IUPAC in plasmids
            plasmids = [sprinkle(rng, p, "NRYKMSW", 0.003) for p in plasmids]
        reads, _ = synth.reads_for(rng, plasmids, 10, topology=topo, frac_full=0.5, frac_partial=0.3,
                                   homopolymer_rate=float(rng.choice([0.0, 1.0, 4.0])),
                                   sub=float(rng.choice([0.01, 0.03, 0.08])), ins=float(rng.choice([0.005, 0.025, 0.06])), dele=float(rng.choice([0.005, 0.025, 0.06])))
        if kind == 2:   # N in reads
            reads = [sprinkle(rng, r, "N", 0.004) for r in reads]
        if kind == 3:   # repeats inside a plasmid (tandem duplication) -> overlapping regions in the query dimension
            p0 = plasmids[0]; a = int(rng.integers(100, len(p0) - 400)); plasmids[0] = p0[:a + 300] + p0[a:a + 300] + p0[a + 300:]
            reads, _ = synth.reads_for(rng, plasmids, 10, topology=topo, frac_full=0.7, frac_partial=0.2)
        if kind == 4:   # very short and very long reads
            reads = reads[:5] + [reads[5][:int(rng.integers(22, 120))], reads[6][:int(rng.integers(120, 400))], (reads[7] * 3)[:int(len(plasmids[0]) * 2.4)], reads[8], reads[9][:25]]
        sc = [(3, 1, 1, -2), (5, 2, 2, -3), (2, 2, 1, -1), (4, 1, 3, -2), (3, 1, 1, -2)][case % 5]
        params = dict(gap_open_penalty=sc[0], gap_extend_penalty=sc[1], match_score=sc[2], mismatch_score=sc[3], topology_of_dna=topo)
        omit = bool(case % 7 != 3)
        want = pool.map(port_read, [(plasmids, r, params, omit) for r in reads], chunksize=1)
        try:
            res = sm.align_all(plasmids, reads, params, omit, engine=eng)
        except Exception as e:
            print("case", case, "ENGINE ERROR", e); nbad += 1; continue
        n2 = 2 * npl
        for qi, w in enumerate(want):
            ntot += 1
            got_status = res.status[qi * n2:(qi + 1) * n2]
            if w == "ASSERT":
                nassert += 1
                if not (got_status == 2).any():
                    print("case", case, "read", qi, "port asserted but engine did not flag"); nbad += 1
                continue
            got = [(res.cigar_string(qi * n2 + j), int(res.score[qi * n2 + j]), res.offset(qi * n2 + j)) for j in range(n2)]
            if got != w:
                nbad += 1
                jbad = [j for j in range(n2) if got[j] != w[j]]
                print("case", case, "kind", kind, "topo", topo, "sc", sc, "read", qi, "len", len(reads[qi]), "pair-strands", jbad, "status", got_status[jbad].tolist(),
                      [(got[j][1], w[j][1], got[j][2], w[j][2], len(got[j][0]), len(w[j][0])) for j in jbad][:3])
    print("reads", ntot, "mismatching", nbad, "port-asserts", nassert)

if __name__ == "__main__":
    main()
