import os, sys
import numpy as np
sys.path.insert(0, ".")
from multiplexnanopore_b200 import Engine
from tests import util
eng = Engine(0)
rng = np.random.default_rng(1)
m, n, cnt = 3000, 4000, 2400
ref = util.rand_seq(rng, n); q = util.rand_seq(rng, m)
pool, s1o, s1l, s2o, s2l = util.pack_problems([(q, ref)])
eng.upload_pool(pool)
rep = lambda a: np.repeat(a, cnt)
eng.align_prepare(rep(s1o), rep(s1l), rep(s2o), rep(s2l), np.ones(cnt, np.uint8), 3, 1, 1, -2)
for it in range(2):
    eng.align_run()
print("ok")
