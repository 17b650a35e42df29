import sys; sys.path.insert(0,'.')
import numpy as np
from multiplexnanopore_b200 import savemoney as sm, synth, Engine
from oracle import oracle as orc, savemoney_port as port
from tests import util
eng=Engine(0)
rng = np.random.default_rng(31)
plasmids = synth.plasmid_set(rng, 4, 1200, 2600, backbone=600)
reads, _ = synth.reads_for(rng, plasmids, 10, topology=0, frac_full=0.6, frac_partial=0.25, homopolymer_rate=1.0)
params = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
rec=[]
orig=port.dp
def dp2(mode, ref, query, sc, log):
    rec.append((mode, query, ref)); return orig(mode, ref, query, sc, log)
port.dp=dp2
qi=1
w = port.align_read(plasmids, reads[qi], params, True)
probs=[(q,r) for _,q,r in rec]; modes=[m for m,_,_ in rec]
pool,a,b,c,d=util.pack_problems(probs)
eng.upload_pool(pool)
res=eng.align_batch(a,b,c,d,np.array(modes,np.uint8),3,1,1,-2)
bad=0
for p,((q,r),m) in enumerate(zip(probs,modes)):
    o=orc.align(m,q,r,3,1,1,-2)
    if res.cigar_string(p)!=o.cigar_rle or res.score[p]!=o.score or res.end_query[p]!=o.end_query or res.end_ref[p]!=o.end_ref:
        bad+=1; print("DP mismatch", p, m, len(q), len(r), res.score[p], o.score, (res.end_query[p],res.end_ref[p]),(o.end_query,o.end_ref))
print("n probs", len(probs), "bad", bad, [ (m,len(q),len(r)) for (q,r),m in zip(probs,modes) if m!=0])
