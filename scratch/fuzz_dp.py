"""one-off randomized DP parity campaign (all four modes, sizes up to 6 k x 6 k, several scorings, team thresholds lowered,
guard bands on, the column threshold of the checkpointed traceback cycling 1 500 / 1 / 200): CUDA engine vs the C oracle.  Run on the GPU box:  python scratch/fuzz_dp.py [seed] [n_rounds]"""
import os, sys, time
sys.path.insert(0, '.')
os.environ["SMX_GUARD"] = "1"
import numpy as np
import multiprocessing as mp
from multiplexnanopore_b200 import Engine
from oracle import oracle as orc
from tests import util

def one(args):
    mode, q, r, sc = args
    o = orc.align(mode, q, r, *sc)
    return (o.score, o.cigar_rle, o.end_query, o.end_ref, o.beg_query, o.beg_ref)

def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    orc.build()
    eng = Engine(0)
    pool = mp.get_context("fork").Pool(processes=os.cpu_count())
    nbad = ntot = 0
    for rd in range(rounds):
        rng = np.random.default_rng(seed0 + rd)
        sc = [(3, 1, 1, -2), (5, 2, 2, -3), (2, 2, 1, -1), (4, 1, 3, -2), (1, 1, 1, -1), (6, 3, 2, -4)][rd % 6]
        if rd % 2:
            os.environ.update(SMX_TEAM2_MCELLS="0.4", SMX_TEAM4_MCELLS="1.0", SMX_TEAM8_MCELLS="2.5")
        else:
            for k in ("SMX_TEAM2_MCELLS", "SMX_TEAM4_MCELLS", "SMX_TEAM8_MCELLS"):
                os.environ.pop(k, None)
        os.environ["SMX_CKPT_MINCOLS"] = ["1500", "1", "200"][rd % 3]   # checkpointed traceback: default, every packed problem of >= 513 rows, most
        probs, modes = [], []
        for _ in range(260):
            m = int(rng.choice([rng.integers(1, 40), rng.integers(33, 300), rng.integers(250, 1100), rng.integers(1000, 6000)], p=[0.2, 0.35, 0.3, 0.15]))
            n = int(rng.choice([rng.integers(1, 60), rng.integers(30, 400), rng.integers(300, 2500), rng.integers(2000, 6000)], p=[0.2, 0.35, 0.3, 0.15]))
            ref = util.rand_seq(rng, n, b"ACGT" if rng.random() < 0.9 else b"ACGTN")
            u = rng.random()
            if u < 0.55:
                src = (ref * (m // n + 2))
                a = int(rng.integers(0, n))
                q = util.mutate(rng, src[a:a + m], float(rng.choice([0.01, 0.05, 0.15])), float(rng.choice([0.0, 0.03, 0.08])), float(rng.choice([0.0, 0.03, 0.08])))[:m] or "A"
            elif u < 0.75:
                unit = util.rand_seq(rng, int(rng.integers(1, 5)))
                q = (unit * (m // len(unit) + 1))[:m]
                ref = (unit * (n // len(unit) + 1))[:n] if rng.random() < 0.5 else ref
            else:
                q = util.rand_seq(rng, m, b"ACGT" if rng.random() < 0.9 else b"ACGTN")
            probs.append((q, ref)); modes.append(int(rng.integers(0, 4)))
        pool_b, s1o, s1l, s2o, s2l = util.pack_problems(probs)
        eng.upload_pool(pool_b)
        t0 = time.time()
        res = eng.align_batch(s1o, s1l, s2o, s2l, np.array(modes, np.uint8), *sc)
        t1 = time.time()
        want = pool.map(one, [(mo, q, r, sc) for (q, r), mo in zip(probs, modes)], chunksize=4)
        for p, w in enumerate(want):
            ntot += 1
            got = (int(res.score[p]), res.cigar_string(p), int(res.end_query[p]), int(res.end_ref[p]), int(res.beg_query[p]), int(res.beg_ref[p]))
            if got != w:
                nbad += 1
                print("round", rd, "sc", sc, "problem", p, "mode", modes[p], "m", len(probs[p][0]), "n", len(probs[p][1]), "score", got[0], w[0], "ends", got[2:], w[2:], "cigar equal", got[1] == w[1])
        print(f"round {rd} scoring {sc}: {len(probs)} problems, engine {t1 - t0:.2f} s, oracle {time.time() - t1:.1f} s, guard errors {eng.guard_errors}", flush=True)
    print("problems", ntot, "mismatching", nbad, "guard errors", eng.guard_errors)

if __name__ == "__main__":
    main()
