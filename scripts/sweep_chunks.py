"""In-process sweep of the chunk size of the exhaustive launch (PIPSORT_EXH_CHUNK is read at every launch; 0 = the
planner's own choice): kernel time (CUDA events around the launch, L2 flushed) per locus size and chunk target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipsort_b200 as P
from pipsort_b200 import synth
sizes = [int(x) for x in sys.argv[1:]] or [60, 100, 150, 200, 300, 600]
for n in sizes:
    L = synth.make_locus(n)
    e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
    for _ in range(5):
        e.reset(); e.run_exhaustive(3)
    e.sync()
    res = {}
    targets = [0, 8, 12, 16, 20, 24, 28, 32, 40, 48, 64, 96, 128, 256]
    if n >= 600:
        targets = [0, 64, 128, 256, 512, 1024, 2048]
    for rnd in range(3):
        for tg in targets:
            if tg:
                os.environ["PIPSORT_EXH_CHUNK"] = str(tg)
            else:
                os.environ.pop("PIPSORT_EXH_CHUNK", None)
            ks = []
            for rep in range(6):
                e.reset(); e.flush_l2(); e.run_exhaustive(3); ks.append(e.last_kernel_ms())
            res.setdefault(tg, []).append(min(ks[1:]))
    os.environ.pop("PIPSORT_EXH_CHUNK", None)
    print(f"n={n} U={L.U}: " + "  ".join(f"{tg}:{1e3 * min(v):.1f}" for tg, v in res.items()) + "  (chunk target: kernel us)", flush=True)
    e.close()
