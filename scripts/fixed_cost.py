"""Kernel time of the exhaustive launch vs the fraction of the rank space it covers (150-SNP locus): fixed vs variable cost."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipsort_b200 as P
from pipsort_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
L = synth.make_locus(n)
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
tot = e.total_ranks(3)
for parts in (1, 2, 4, 8, 16, 32, 64):
    b = e.shard_ranks(3, parts)
    for which in (0, parts - 1):
        ks = []
        for rep in range(6):
            e.reset()
            if len(sys.argv) > 2: e.flush_l2()
            e.run_exhaustive(3, b[which], b[which + 1]); ks.append(e.last_kernel_ms())
        print(f"1/{parts:<3d} shard {which:2d}: ranks {b[which + 1] - b[which]:8d}  kernel us min {1e3 * min(ks[1:]):7.1f} mean {1e3 * sum(ks[1:]) / 5:7.1f}", flush=True)
e.close()
