"""Per-warp time stamps of the exhaustive kernel on a synthetic locus (needs the tracing build var_trace.so: a scratch copy
of csrc with %globaltimer stamps -- 0 kernel entry, 1 descriptor loaded, 2 a values, 3 window table, 4 lane values + first
load issued, 5 last step done, 6 chunk flushed; word 7 = the descriptor's step count).
python scripts/trace_chunks.py [n_snps] [world rank]"""
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
names = ["descriptor", "a values", "window table", "lane values", "stepping", "flush"]
for it in range(4):
    buf = np.zeros(NW * 8, dtype=np.uint64)
    e.reset(); e.flush_l2(); e.run_exhaustive(3, b[rank], b[rank + 1]); e.sync()
    lib.pipsort_debug_trace(buf.ctypes.data_as(C.c_void_p), NW * 8)
    t = buf.reshape(NW, 8).astype(np.int64)
    t0 = t[:, 0].min()
    ok = (t[:, 6] > t0) & (t[:, 1] > t0)          # warps that took a size-3 chunk in THIS pass
    rel = (t[ok, :7] - t0) / 1e3
    steps = (t[ok, 7] & 0x0fffffff)
    med = [float(np.median(rel[:, i + 1] - rel[:, i])) for i in range(6)]
    print("pass %d kernel_ms %.4f warps %d steps/warp %.1f | " % (it, e.last_kernel_ms(), ok.sum(), steps.mean()) +
          " | ".join("%s +%.2f" % (nm, m) for nm, m in zip(names, med)) + " | last chunk end %.1f us" % rel[:, 6].max())
