import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipsort_b200 as P
from pipsort_b200 import synth
L = synth.make_locus(150)
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
tot = e.total_ranks(3)
e.reset(); e.run_exhaustive(3); e.sync()
for name, lo, hi in [("null only", 0, 1), ("singles", 0, 181), ("<=pairs", 0, 181 + 16110), ("first 32 triples", 16291, 16291 + 32), ("first 5000 triples", 16291, 21291),
                     ("last 5000 triples", tot - 5000, tot), ("all", 0, tot)]:
    ks = []
    for rep in range(8):
        e.reset(); e.run_exhaustive(3, lo, hi); ks.append(e.last_kernel_ms())
    t = []
    for rep in range(8):
        e.reset(); e.sync(); e.timer_begin(); e.run_exhaustive(3, lo, hi); t.append(e.timer_end())
    print(f"{name:20s} kernel-event us min {1e3 * min(ks[1:]):6.1f}   begin..end us min {1e3 * min(t[1:]):6.1f}", flush=True)
e.close()
