import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PIPSORT_TRACE"] = "1"
import numpy as np, torch
import pipsort_b200 as P
from pipsort_b200 import synth
L = synth.make_locus(150)
sig = torch.from_numpy(np.concatenate([s.ravel() for s in L.sigma])).pin_memory().numpy()
z = torch.from_numpy(np.concatenate(L.z)).pin_memory().numpy()
for it in range(6):
    t = time.perf_counter()
    r = P.posterior_exhaustive(L.num_snps, sig, z, L.d, L.K, L.snp_map, 3, gamma=L.gamma, sharing_param=L.sharing_param)
    print(f"python total {1e6 * (time.perf_counter() - t):.1f} us", flush=True)
