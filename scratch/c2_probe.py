import sys, time, json; sys.path.insert(0,'.')
import numpy as np
from multiplexnanopore_b200 import savemoney as sm, synth, Engine
eng=Engine(0)
n=int(sys.argv[1]) if len(sys.argv)>1 else 500
t=time.time(); plasmids, reads, truth = synth.config_c2(n_reads=n); print("gen", time.time()-t, [len(p) for p in plasmids])
params = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
for it in range(2):
    t=time.time(); res = sm.align_all(plasmids, reads, params, True, engine=eng); dt=time.time()-t
    print("wall", dt, "reads/s", n/dt); print({k:(round(v,4) if v<1e6 else v) for k,v in res.stats.items()})
# assignment check
sc = res.score.reshape(n, -1).astype(float)
norm = sc / np.repeat([len(p) for p in plasmids], 2)[None,:]
best = norm.argmax(1)
ok = sum(1 for i,(pi,rc) in enumerate(truth) if best[i]==2*pi+int(rc))
print("argmax==truth", ok, "/", n, "status counts", np.bincount(res.status, minlength=3))
