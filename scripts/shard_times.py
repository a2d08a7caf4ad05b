"""Kernel time of every shard of an N-way split, run one after the other on ONE GPU (how balanced is the shard planner?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipsort_b200 as P
from pipsort_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = synth.make_locus(n)
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
e.reset(); e.run_exhaustive(3); full = e.last_kernel_ms()
b = e.shard_ranks(3, parts)
ts = []
for i in range(parts):
    ks = []
    for rep in range(3):
        e.reset(); e.run_exhaustive(3, b[i], b[i + 1]); ks.append(e.last_kernel_ms())
    ts.append(min(ks))
print(f"n={n} parts={parts} full {full:.3f} ms; shards ms: " + " ".join(f"{t:.3f}" for t in ts) + f"; max/mean {max(ts) / (sum(ts) / parts):.3f}, sum/full {sum(ts) / full:.3f}")
e.close()
