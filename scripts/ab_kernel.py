"""A/B of builds of the library on the exhaustive kernel (L2 flushed before every launch, CUDA events around the kernel):
python scripts/ab_kernel.py lib1.so lib2.so ...  -- each build in its own process, interleaved rounds."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r)
import pipsort_b200 as P
from pipsort_b200 import synth
import numpy as np
for n, reps in ((150, 40), (1500, 3)):
    L = synth.make_locus(n)
    e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
    for _ in range(3):
        e.reset(); e.run_exhaustive(3)
    ks = []
    for _ in range(reps):
        e.reset(); e.flush_l2(); e.run_exhaustive(3); ks.append(e.last_kernel_ms())
    print(n, "median %%.4f min %%.4f" %% (float(np.median(ks)), min(ks)), flush=True)
    e.close()
''' % ROOT
libs = sys.argv[1:]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ, PIPSORT_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True).stdout.strip().replace("\n", " | ")
        print(os.path.basename(lib), out, flush=True)
