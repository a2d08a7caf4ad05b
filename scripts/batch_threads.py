import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pipsort_b200 as P
from pipsort_b200 import synth
L = synth.make_locus(150)
sig = torch.from_numpy(np.concatenate([s.ravel() for s in L.sigma])).pin_memory(); z = torch.from_numpy(np.concatenate(L.z)).pin_memory()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 160
loci = []
for i in range(nb):
    sg = sig.clone().pin_memory(); zz = z.clone().pin_memory()
    loci.append(dict(num_snps=L.num_snps.copy(), sigma=sg.numpy(), z=zz.numpy(), d=L.d.copy(), K=L.K, snp_map=L.snp_map.copy(), gamma=L.gamma, sharing_param=L.sharing_param, _k=(sg, zz)))
P.posterior_exhaustive_batch(loci[:16], 3)
b = P.LocusBatch(loci, 3)
for rep in range(4):
    torch.cuda.synchronize(); t = time.perf_counter(); b.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(os.environ.get("PIPSORT_BATCH_THREADS", "default"), "ms per locus %.4f" % (1e3 * dt / nb), flush=True)
rs = b.results()
assert all(r.n_configs == 12197751 for r in rs)
