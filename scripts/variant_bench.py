"""Times the exhaustive kernel of several builds (PIPSORT_B200_LIB) on the B150c3 / B1500c3 loci."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import pipsort_b200 as P
from pipsort_b200 import synth
for n, reps in ((150, 20), (1500, 3)):
    L = synth.make_locus(n)
    e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
    for _ in range(3):
        e.reset(); e.run_exhaustive(3)
    e.sync()
    ks = []
    for _ in range(reps):
        e.reset(); e.run_exhaustive(3); ks.append(e.last_kernel_ms())
    print(n, "kernel ms min %%.4f mean %%.4f" %% (min(ks), sum(ks) / len(ks)), flush=True)
    e.close()
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, PIPSORT_B200_LIB=os.path.abspath(lib))
    print("==", lib, flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
