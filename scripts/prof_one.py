"""One locus, a few exhaustive passes: the target of the ncu captures under profiles/ (python scripts/prof_one.py 1500 3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipsort_b200 as P
from pipsort_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = synth.make_locus(n)
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=c)
ks = []
for _ in range(reps):
    e.reset(); e.run_exhaustive(c); ks.append(e.last_kernel_ms())
r = e.read()
print(n, c, "kernel ms", ["%.4f" % k for k in ks], "total", r.total, flush=True)
e.close()
