"""One neighbourhood of the stochastic shotgun search at the size of BASELINE.json configs[4] (5000 SNPs/study, c=5): the
target of the ncu capture of score_batch_kernel, and the timing behind the LD-gather GB/s figure."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pipsort_b200 as P
from pipsort_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
c = 5
L = synth.make_locus(n, overlap=0.8)
U = L.U
cur = sorted({int(np.where(L.snp_map[0] == int(np.argmax(np.abs(L.z[0]))))[0][0]), 17, U // 2 - 1, U - 2})
non = [g for g in range(U) if g not in set(cur)]
minus = [cur[:m] + cur[m + 1:] for m in range(len(cur))]
nbd = [sorted(m + [g]) for g in non for m in minus] + minus + [sorted(cur + [g]) for g in non]
idx = np.full((len(nbd), c), -1, dtype=np.int32)
for i, v in enumerate(nbd):
    idx[i, :len(v)] = v
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=c)
import torch
d_idx = torch.from_numpy(idx).cuda()
d_out = torch.empty(len(nbd), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ms = []
flush = not os.environ.get("PROF_NO_FLUSH")      # the 400 MB of LD exceed L2 (126 MB) either way; the flush also evicts the kernel's code
for rep in range(8):
    e.reset()
    if flush:
        e.flush_l2()
    e.sync()
    e.timer_begin()
    e.score_union_configs_device(d_idx.data_ptr(), len(nbd), c, 0, d_out.data_ptr())
    ms.append(e.timer_end())
r = e.read()
# LD gather: per union configuration of k SNPs the kernel reads the k x k causal sub-blocks of both studies (8-byte words)
present = [(L.snp_map[0][v] >= 0, L.snp_map[1][v] >= 0) for v in map(np.array, nbd)]
gather_bytes = sum(8 * (int(p0.sum()) ** 2 + int(p1.sum()) ** 2) for p0, p1 in present)
best = min(ms)
print(f"n={n} U={U} neighbourhood={len(nbd)} union configs, {r.n_configs} expanded configurations")
print("kernel ms:", " ".join(f"{x:.3f}" for x in ms))
print(f"{r.n_configs / (best * 1e-3):.4e} configs/s; LD gather {gather_bytes} B -> {gather_bytes / (best * 1e-3) / 1e9:.3f} GB/s")
e.close()
