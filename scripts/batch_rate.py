"""Per-locus time of the pipelined batch call for several locus sizes (host-bound or device-bound?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pipsort_b200 as P
from pipsort_b200 import synth
for n in (20, 40, 90, 150, 300):
    L = synth.make_locus(n)
    sig = torch.from_numpy(np.concatenate([s.ravel() for s in L.sigma])).pin_memory().numpy()
    z = torch.from_numpy(np.concatenate(L.z)).pin_memory().numpy()
    locus = dict(num_snps=L.num_snps, sigma=sig, z=z, d=L.d, K=L.K, snp_map=L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param)
    P.posterior_exhaustive_batch([locus] * 8, 3)
    best = 1e9
    for rep in range(3):
        t = time.perf_counter(); P.posterior_exhaustive_batch([locus] * 100, 3); best = min(best, time.perf_counter() - t)
    e = P.Engine(L.num_snps, sig, z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
    e.reset(); e.run_exhaustive(3); e.reset(); e.run_exhaustive(3); k = e.last_kernel_ms(); e.close()
    print(f"n={n}: batch {1e3 * best / 100 * 1e3:.1f} us per locus, exhaustive kernel alone {1e3 * k:.1f} us", flush=True)
