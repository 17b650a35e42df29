import sys; sys.path.insert(0,'.')
import numpy as np, ctypes
from multiplexnanopore_b200 import Engine
from oracle import oracle as orc
from tests import util
eng=Engine(0)
q,ref='AT','CACT'
for mode in (3,0):
    pool,a,b,c,d=util.pack_problems([(q,ref)])
    eng.upload_pool(pool)
    r=eng.align_batch(a,b,c,d,[mode],1,1,3,-1)
    print(mode, r.score, r.cigar_string(0))
    n=len(ref); nb=(n+31)*32
    buf=np.zeros(nb,np.uint8)
    eng._lib.smx_debug_copy_trace(eng._h, buf.ctypes.data_as(ctypes.c_void_p), nb)
    t=buf.reshape(n+31,32)
    for i in range(len(q)):
        print([int(t[j+i,i]) for j in range(n)])
