import sys
import numpy as np
sys.path.insert(0, ".")
from multiplexnanopore_b200 import Engine
from oracle import oracle as orc
from tests import util
orc.build()
eng = Engine(0)
rng = np.random.default_rng(12)
probs, modes = [], []
for m in (257, 300, 511, 512, 513, 700, 1100):
    for n in (1, 2, 31, 32, 33, 64, 100, 257, 600):
        ref = util.rand_seq(rng, n)
        q = util.mutate(rng, (ref * (m // n + 2))[:m], 0.1, 0.0, 0.0)[:m]
        for mode in range(3):
            probs.append((q, ref)); modes.append(mode)
go, ge, ma, mi = 3, 1, 1, -2
pool, s1o, s1l, s2o, s2l = util.pack_problems(probs)
eng.upload_pool(pool)
res = eng.align_batch(s1o, s1l, s2o, s2l, np.asarray(modes, np.uint8), go, ge, ma, mi)
bad = 0
for p, ((q, r), mode) in enumerate(zip(probs, modes)):
    o = orc.align(int(mode), q, r, go, ge, ma, mi)
    ok_s = res.score[p] == o.score
    ok_e = (res.end_query[p], res.end_ref[p]) == (o.end_query, o.end_ref)
    ok_c = res.cigar_string(p) == o.cigar_rle
    if not (ok_s and ok_e and ok_c):
        bad += 1
        if bad <= 40:
            print(f"p{p} mode {mode} m={len(q)} n={len(r)}: score {res.score[p]} vs {o.score}; end {(res.end_query[p], res.end_ref[p])} vs {(o.end_query, o.end_ref)}; cigar_ok {ok_c}")
print("bad", bad, "of", len(probs))
for p in (14, 44):
    q, r = probs[p]
    o = orc.align(int(modes[p]), q, r, go, ge, ma, mi)
    print(p, "GPU ", res.cigar_string(p))
    print(p, "ORC ", o.cigar_rle)
