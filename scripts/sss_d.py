"""Config D (BASELINE.json configs[4]): 5000 SNPs/study, c=5 stochastic shotgun search; times create (raw LD, on-device
pre-processing vs pre-processed input), one neighbourhood launch, and a bounded number of search rounds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pipsort_b200 as P
from pipsort_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
c = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t = time.time(); L = synth.make_locus(n, overlap=0.8); print(f"synth {time.time() - t:.1f}s U={L.U}", flush=True)
t = time.time()
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=c)
e.sync(); print(f"create (pre-processed input) {time.time() - t:.3f}s", flush=True)
for rep in range(2):
    t = time.time(); r, it, why = e.sss(c, max_iterations=iters); dt = time.time() - t
    print(f"sss rep{rep}: {it} rounds in {dt:.3f}s = {1e3 * dt / max(it, 1):.2f} ms/round, {r.n_configs} configurations, "
          f"{r.n_configs / dt:.3e} configs/s, stop={why}, total={r.total:.6f}", flush=True)
e.close()
if len(sys.argv) > 4:
    t = time.time()
    e = P.Engine(L.num_snps, L.sigma, L.z, L.d, 0.0, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=c, raw_ld=True)
    e.sync(); print(f"create (raw LD, on-device PSD shift + eigen) {time.time() - t:.3f}s", [e.prep_info(s) for s in range(2)], flush=True)
    e.close()
