"""Where the time of the 5000-SNP pre-processing goes: one LU, one Cholesky, one symmetric eigen-decomposition (cuSOLVER
behind torch.linalg, float64) at n SNPs, and the create() phases with pre-processed input (PIPSORT_TRACE_CREATE=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pipsort_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
L = synth.make_locus(n, overlap=0.8)
A = torch.from_numpy(L.sigma[0]).cuda()
torch.cuda.synchronize()


def timed(name, f, reps=3):
    for i in range(reps):
        torch.cuda.synchronize(); t = time.time(); r = f(); torch.cuda.synchronize()
        print(f"{name} rep{i}: {1e3 * (time.time() - t):.1f} ms", flush=True)
    return r


timed("lu_factor", lambda: torch.linalg.lu_factor(A + 0.05 * torch.eye(n, dtype=A.dtype, device="cuda")))
timed("cholesky", lambda: torch.linalg.cholesky(A + 0.05 * torch.eye(n, dtype=A.dtype, device="cuda")))
timed("eigvalsh", lambda: torch.linalg.eigvalsh(A), reps=2)
timed("eigh", lambda: torch.linalg.eigh(A), reps=2)
# two at once on two streams (do independent factorizations overlap?)
B = torch.from_numpy(L.sigma[1]).cuda()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.time()
with torch.cuda.stream(s1):
    torch.linalg.eigh(A)
with torch.cuda.stream(s2):
    torch.linalg.eigh(B)
torch.cuda.synchronize(); print(f"two eigh on two streams: {1e3 * (time.time() - t):.1f} ms", flush=True)
torch.cuda.synchronize(); t = time.time()
with torch.cuda.stream(s1):
    for k in range(4):
        torch.linalg.lu_factor(A + 0.01 * k * torch.eye(n, dtype=A.dtype, device="cuda"))
with torch.cuda.stream(s2):
    for k in range(4):
        torch.linalg.lu_factor(B + 0.01 * k * torch.eye(n, dtype=A.dtype, device="cuda"))
torch.cuda.synchronize(); print(f"2 x 4 lu_factor on two streams: {1e3 * (time.time() - t):.1f} ms", flush=True)
w = torch.linalg.eigvalsh(A).cpu().numpy()
for k in (0, 1, 5, 10, 20, 25, 26, 27, 30):
    a = 0.01 * k
    print(f"shift {a:.2f}: min eig {w.min() + a:.3e}, log2 det {np.sum(np.log2(np.abs(w + a))):.1f}, negatives {(w + a < 0).sum()}")
del A, B
torch.cuda.empty_cache()
import pipsort_b200 as P
os.environ["PIPSORT_TRACE_CREATE"] = "1"
for rep in range(2):
    t = time.time()
    e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=5)
    t1 = time.time(); e.sync(); t2 = time.time()
    print(f"create rep{rep}: python+C {t1 - t:.3f}s, sync {t2 - t1:.3f}s", flush=True)
    e.close()
