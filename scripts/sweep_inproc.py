"""In-process sweep of the work-queue granularity (PIPSORT_EXH_BW / PIPSORT_EXH_XCH are read at every launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipsort_b200 as P
from pipsort_b200 import synth
sizes = [int(x) for x in sys.argv[1:]] or [60, 90, 120, 150, 200, 300, 450, 600]
for n in sizes:
    L = synth.make_locus(n)
    e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
    for _ in range(5):
        e.reset(); e.run_exhaustive(3)
    e.sync()
    res = {}
    for rnd in range(3):
        for bw, xch in [(0, 0), (32, 1), (32, 2), (32, 4), (16, 1), (8, 1)]:
            os.environ["PIPSORT_EXH_BW"] = str(bw); os.environ["PIPSORT_EXH_XCH"] = str(xch)
            ks = []
            for rep in range(6):
                e.reset(); e.run_exhaustive(3); ks.append(e.last_kernel_ms())
            res.setdefault((bw, xch), []).append(min(ks[1:]))      # (0, 0) = the engine's own choice
    print(f"n={n} U={L.U}: " + "  ".join(f"({bw},{xch}) {1e3 * min(v):.1f}" for (bw, xch), v in res.items()), flush=True)
    e.close()
