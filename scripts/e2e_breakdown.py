import time, sys, numpy as np
sys.path.insert(0,'/root/repo')
import pipsort_b200 as P
from pipsort_b200 import synth
import torch
for n in (150,1500):
    L=synth.make_locus(n)
    sig=np.concatenate([s.ravel() for s in L.sigma]); z=np.concatenate(L.z)
    for it in range(4):
        t0=time.perf_counter()
        e=P.Engine(L.num_snps,sig,z,L.d,L.K,L.snp_map,gamma=L.gamma,sharing_param=L.sharing_param,max_causal=3)
        t1=time.perf_counter()
        e.run_exhaustive(3); t2=time.perf_counter()
        e.sync(); t3=time.perf_counter()
        r=e.read(); t4=time.perf_counter()
        e.close(); t5=time.perf_counter()
        print(n,it,'create %.3f run-call %.3f sync %.3f read %.3f close %.3f ms'%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3))
