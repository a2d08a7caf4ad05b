"""Per-warp time stamps of the exhaustive kernel on the 150-SNP locus (needs the tracing build: a scratch copy of csrc with
%globaltimer stamps at kernel entry, after the descriptor load, before the first step, after the last step, end of chunk,
end of the warp's work).  python scripts/trace_chunks.py [n_snps] [world rank]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PIPSORT_B200_LIB", os.path.join(ROOT, "pipsort_b200", "lib", "var_trace.so"))
import numpy as np
import pipsort_b200 as P
from pipsort_b200 import synth, engine as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
world, rank = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 0)
L = synth.make_locus(n)
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=3)
b = e.shard_ranks(3, world)
lib = E.lib()
lib.pipsort_debug_trace.argtypes = [C.c_void_p, C.c_int]
NW = 148 * 12
for it in range(4):
    e.reset(); e.flush_l2(); e.run_exhaustive(3, b[rank], b[rank + 1]); e.sync()
    buf = np.zeros(NW * 8, dtype=np.uint64)
    lib.pipsort_debug_trace(buf.ctypes.data_as(C.c_void_p), NW * 8)
    t = buf.reshape(NW, 8).astype(np.int64)
    ok = t[:, 4] > 0
    t0 = t[ok, 0].min()
    rel = (t[ok, :6] - t0) / 1e3
    steps = (t[ok, 7] & 0x0fffffff)
    print("pass %d kernel_ms %.4f warps %d | entry %.1f..%.1f | descriptor +%.2f | to first step +%.2f | stepping %.2f (steps %.1f, segs %.1f, %.2f us/step) | flush +%.2f | end of first chunk %.1f..%.1f max-warp-end %.1f"
          % (it, e.last_kernel_ms(), ok.sum(), rel[:, 0].min(), rel[:, 0].max(), np.median(rel[:, 1] - rel[:, 0]), np.median(rel[:, 2] - rel[:, 1]),
             np.median(rel[:, 3] - rel[:, 2]), steps.mean(), t[ok, 6].mean(), np.median((rel[:, 3] - rel[:, 2]) / np.maximum(steps, 1)),
             np.median(rel[:, 4] - rel[:, 3]), rel[:, 4].min(), rel[:, 4].max(), rel[:, 5].max()))
    if it == 3:
        order = np.argsort(rel[:, 4])
        for i in list(order[:3]) + list(order[-5:]):
            print("   warp %4d: entry %.1f desc %.1f first-step %.1f last-step %.1f end %.1f  steps %d segs %d" % (i, *rel[i, :5], steps[i], t[ok, 6][i]))
