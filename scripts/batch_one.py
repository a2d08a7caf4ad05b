"""Eight 150-SNP loci through the pipelined batch call: the target of the launch-list capture of the end-to-end path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pipsort_b200 as P
from pipsort_b200 import synth
L = synth.make_locus(150)
locus = dict(num_snps=L.num_snps, sigma=np.concatenate([s.ravel() for s in L.sigma]), z=np.concatenate(L.z), d=L.d, K=L.K,
             snp_map=L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param)
rs = P.posterior_exhaustive_batch([locus] * 8, 3)
print(rs[0].total, rs[-1].n_configs)
