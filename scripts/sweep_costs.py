"""Calibrates the step-cost constants of the exhaustive kernel's work plan (ExhCostModel, exh_plan.h) on the 150-SNP locus,
whose kernel time is the time of its longest chunk: python scripts/sweep_costs.py  -> kernel ms per (c_chain, c_plain)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys
sys.path.insert(0, %r)
import pipsort_b200 as P
from pipsort_b200 import synth
import numpy as np
L = synth.make_locus(150)
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
for _ in range(3):
    e.reset(); e.run_exhaustive(3)
ks = []
for _ in range(60):
    e.reset(); e.flush_l2(); e.run_exhaustive(3); ks.append(e.last_kernel_ms())
print("%%.4f %%.4f %%.4f" %% (float(np.mean(ks)), float(np.median(ks)), min(ks)))
''' % ROOT
if "--setup" in sys.argv:      # set-up costs per segment / window / a / chunk, in generic warp-steps (PIPSORT_EXH_SETUP)
    for setup in ("0.6,0.9,0.5,0.8", "0.6,1.2,1.6,2.4", "0.6,1.0,1.0,1.5", "1.0,1.5,2.0,2.4", "0.3,0.5,0.3,0.8", "0.6,2.0,2.5,2.4"):
        for rep in range(2):
            env = dict(os.environ, PIPSORT_EXH_SETUP=setup)
            out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
            print("seg,win,a,chunk %s  kernel ms (mean, median, min): %s" % (setup, out.stdout.strip() or out.stderr[-300:]), flush=True)
    sys.exit(0)
grid = [(c, p) for c in (0.3, 0.4, 0.5, 0.6, 0.7) for p in (0.15, 0.25, 0.4)] + [(1.0, 1.0)]
for c, p in grid:
    env = dict(os.environ, PIPSORT_EXH_COSTS="%g,%g,0.45,0.3" % (c, p))
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("c_chain %.2f c_plain %.2f  kernel ms (mean, median, min): %s" % (c, p, out.stdout.strip() or out.stderr[-300:]), flush=True)
