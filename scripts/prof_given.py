"""Explicit-configuration path (-b/-d/-e): throughput of given_configs_kernel on device-resident rows (300-SNP/study locus)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pipsort_b200 as P
from pipsort_b200 import synth
from pipsort_b200.engine import lib, _check
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
groups, kmax = 6, 6
L = synth.make_locus(n)
rng = np.random.default_rng(3)
# rows of up to 6 causal SNPs: sort random draws, drop duplicates by re-drawing whole rows (rare at N = 600)
M = np.sort(rng.integers(0, L.N, size=(rows, groups), dtype=np.int64), axis=1)
dup = (np.diff(M, axis=1) == 0).any(axis=1)
M[dup] = np.arange(groups)[None, :] * 7 + 1
k = rng.integers(0, kmax + 1, size=rows)
M[np.arange(groups)[None, :] >= k[:, None]] = -1
M = np.where(M < 0, -1, M).astype(np.int16)
# negative entries must not break the increasing order of the others: they are skipped by the kernel
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
d = torch.from_numpy(M).cuda()
torch.cuda.synchronize()
ms = []
for rep in range(4):
    e.reset(); e.sync(); e.timer_begin()
    _check(lib().pipsort_score_given_configs_device(e._h, C.c_void_p(d.data_ptr()), rows, groups))
    ms.append(e.timer_end())
r = e.read()
print(f"n={n} rows={rows} groups={groups}: kernel ms " + " ".join(f"{x:.3f}" for x in ms) + f" -> {rows / (min(ms) * 1e-3):.3e} configurations/s; counted {r.n_configs}")
# host-buffer path (H2D of the int16 matrix inside)
import time
t = time.perf_counter(); e.reset(); e.score_given_configs(M); e.sync(); dt = time.perf_counter() - t
print(f"host-buffer call: {dt * 1e3:.2f} ms for {M.nbytes / 1e6:.1f} MB of rows -> {rows / dt:.3e} configurations/s")
e.close()
