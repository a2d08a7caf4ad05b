"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under pipsort_b200/ imports it.

The small file readers here restate /root/reference/util.cpp:86-159 and model.h:86-144 (LD file =
whitespace separated doubles, z file = "name z" per line, snp_map = "rsid,idx0,idx1").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u64 = C.c_uint64


def build(force: bool = False) -> str:
    """Compile oracle/_build/liboracle.so (and, when /root/reference is present, oracle/_ref)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_ncr.restype = _u64
        _lib.oracle_total_union_subsets.restype = _u64
        _lib.oracle_ll_dense.restype = C.c_double
        _lib.oracle_f_block.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


@dataclass
class Locus:
    """Engine-level description of a locus (what crosses the drop-in boundary)."""

    n_snps: np.ndarray            # int32[S]
    sigma: list                   # per study: effective LD  B^T B, float64[n,n]
    z: list                       # per study: B^T S', float64[n]
    K: float                      # S'^T S'
    d: np.ndarray                 # float64[S]   s^2 n_s/min(n) + t^2
    snp_map: np.ndarray           # int32[S,U]   study index or -1
    gamma: float = 0.01
    p: float = 0.75
    names: list = field(default_factory=list)        # per study SNP names
    union_names: list = field(default_factory=list)  # rsid per union position
    add_diag: list = field(default_factory=list)

    @property
    def S(self):
        return len(self.n_snps)

    @property
    def U(self):
        return self.snp_map.shape[1]

    @property
    def N(self):
        return int(self.n_snps.sum())

    def flat(self):
        sig = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64).ravel() for s in self.sigma]))
        z = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64) for v in self.z]))
        return (np.ascontiguousarray(self.n_snps, dtype=np.int32), sig, z,
                np.ascontiguousarray(self.d, dtype=np.float64), np.ascontiguousarray(self.snp_map, dtype=np.int32))


def d_per_study(sample_sizes, s_squared=5.2, t_squared=0.52):
    """postcal.cpp:66,89: s^2 * double(n_s) / int(min n) + t^2."""
    mn = int(min(sample_sizes))
    return np.array([s_squared * (float(n) / mn) + t_squared for n in sample_sizes], dtype=np.float64)


def read_ld(path):
    vals = []
    with open(path) as f:                      # util.cpp:86-96 stops at the first non-numeric token
        for tok in f.read().split():
            try:
                vals.append(float(tok))
            except ValueError:
                break
    n = int(np.sqrt(len(vals)))                # model.h:98
    return np.array(vals[: n * n], dtype=np.float64).reshape(n, n)


def read_z(path):
    names, z = [], []
    with open(path) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            names.append(parts[0])
            z.append(float(parts[1]))
    return names, np.array(z, dtype=np.float64)


def read_snp_map(path, S=2):
    names, cols = [], [[] for _ in range(S)]
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            parts = line.split(",")
            names.append(parts[0])
            for s in range(S):
                cols[s].append(int(parts[1 + s]))
    return names, np.array(cols, dtype=np.int32)


def preprocess(ld, z):
    """model.h:171-264 for one study -> (sigma_eff, z_eff, K, add_diag, B, Sprime)."""
    n = ld.shape[0]
    ld = np.ascontiguousarray(ld, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    sig = np.empty((n, n)); ze = np.empty(n); B = np.empty((n, n)); Sp = np.empty(n)
    K = C.c_double(); add = C.c_double()
    rc = lib().oracle_preprocess(n, _d(ld), _d(z), _d(sig), _d(ze), C.byref(K), C.byref(add), _d(B), _d(Sp))
    assert rc == 0
    return sig, ze, K.value, add.value, B, Sp


def load_locus(ld_files, z_files, snp_map_file, sample_sizes, gamma=0.01, p=0.75, s_squared=5.2, t_squared=0.52,
               raw=False):
    """Files -> Locus, through the oracle's own restatement of the host pre-processing.

    raw=True skips the PSD/eigen step (sigma = LD as read, K from a solve) -- for synthetic loci that
    are positive definite by construction."""
    sig, zs, names, adds = [], [], [], []
    K = 0.0
    for lf, zf in zip(ld_files, z_files):
        ld = read_ld(lf)
        nm, z = read_z(zf)
        assert ld.shape[0] == len(nm)
        if raw:
            se, ze, k, a = ld, z, float(z @ np.linalg.solve(ld, z)), 0.0
        else:
            se, ze, k, a, _, _ = preprocess(ld, z)
        sig.append(se); zs.append(ze); names.append(nm); adds.append(a)
        K += k
    union_names, smap = read_snp_map(snp_map_file, len(ld_files))
    return Locus(n_snps=np.array([s.shape[0] for s in sig], dtype=np.int32), sigma=sig, z=zs, K=K,
                 d=d_per_study(sample_sizes, s_squared, t_squared), snp_map=smap, gamma=gamma, p=p, names=names,
                 union_names=union_names, add_diag=adds)


@dataclass
class Result:
    total: float
    post: np.ndarray
    noCausal: np.ndarray
    sharedPips: np.ndarray
    sharedLL: np.ndarray
    notSharedLL: np.ndarray
    n_eval: int = 0
    extra: dict = field(default_factory=dict)

    def pips(self):
        """postcal.h:277-283 special_exp against total."""
        return np.where(self.post == 0, 0.0, np.exp(self.post - self.total))

    def shared_pips(self):
        return np.where(self.sharedPips == 0, 0.0, np.exp(self.sharedPips - self.total))

    def no_causal(self):
        return np.where(self.noCausal == 0, 0.0, np.exp(self.noCausal - self.total))


def total_union_subsets(U, c):
    return int(lib().oracle_total_union_subsets(int(U), int(c)))


def exhaustive(L: Locus, c: int, rank_begin: int = 0, rank_end: int | None = None) -> Result:
    n, sig, z, d, smap = L.flat()
    if rank_end is None:
        rank_end = total_union_subsets(L.U, c)
    post = np.zeros(L.N); nc = np.zeros(L.S); sp = np.zeros(L.U); sl = np.zeros(L.U); nl = np.zeros(L.U)
    tot = C.c_double(); ne = _u64()
    rc = lib().oracle_exhaustive(L.S, _i(n), _d(sig), _d(z), C.c_double(L.K), _d(d), L.U, _i(smap), int(c),
                                 C.c_double(L.gamma), C.c_double(L.p), _u64(rank_begin), _u64(rank_end),
                                 C.byref(tot), _d(post), _d(nc), _d(sp), _d(sl), _d(nl), C.byref(ne))
    assert rc == 0
    return Result(tot.value, post, nc, sp, sl, nl, int(ne.value))


def exhaustive_omp(L: Locus, c: int, rank_begin: int, rank_end: int, threads: int = 0):
    n, sig, z, d, smap = L.flat()
    tot = C.c_double(); ne = _u64()
    rc = lib().oracle_exhaustive_omp(L.S, _i(n), _d(sig), _d(z), C.c_double(L.K), _d(d), L.U, _i(smap), int(c),
                                     C.c_double(L.gamma), C.c_double(L.p), _u64(rank_begin), _u64(rank_end),
                                     int(threads), C.byref(tot), C.byref(ne))
    assert rc == 0
    return tot.value, int(ne.value)


def exhaustive_omp_full(L: Locus, c: int, rank_begin: int = 0, rank_end: int | None = None, threads: int = 0) -> Result:
    """exhaustive() on all host cores (per-thread accumulators merged in log space): every accumulator, for the sizes the
    sequential walk would take minutes on (12 M configurations of the 150-SNP c=3 locus, rank ranges of the 1500-SNP one)."""
    n, sig, z, d, smap = L.flat()
    if rank_end is None:
        rank_end = total_union_subsets(L.U, c)
    post = np.zeros(L.N); nc = np.zeros(L.S); sp = np.zeros(L.U); sl = np.zeros(L.U); nl = np.zeros(L.U)
    tot = C.c_double(); ne = _u64()
    rc = lib().oracle_exhaustive_omp_full(L.S, _i(n), _d(sig), _d(z), C.c_double(L.K), _d(d), L.U, _i(smap), int(c),
                                          C.c_double(L.gamma), C.c_double(L.p), _u64(rank_begin), _u64(rank_end),
                                          int(threads), C.byref(tot), _d(post), _d(nc), _d(sp), _d(sl), _d(nl), C.byref(ne))
    assert rc == 0
    return Result(tot.value, post, nc, sp, sl, nl, int(ne.value))


def score_union_configs(L: Locus, idx: np.ndarray, make_updates=None, state: Result | None = None):
    """sss_postcal.cpp:447-685 for a batch; returns (max_abs_l[n], Result with updated accumulators)."""
    n, sig, z, d, smap = L.flat()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    nn, kmax = idx.shape
    if state is None:
        state = Result(0.0, np.zeros(L.N), np.zeros(L.S), np.zeros(L.U), np.zeros(L.U), np.zeros(L.U))
    post = state.post.copy(); nc = state.noCausal.copy(); sp = state.sharedPips.copy()
    sl = state.sharedLL.copy(); nl = state.notSharedLL.copy()
    tot = C.c_double(state.total)
    out = np.zeros(nn)
    mu = None
    if make_updates is not None:
        mu = np.ascontiguousarray(make_updates, dtype=np.uint8)
    rc = lib().oracle_score_union_configs(L.S, _i(n), _d(sig), _d(z), C.c_double(L.K), _d(d), L.U, _i(smap),
                                          C.c_double(L.gamma), C.c_double(L.p), _i(idx), nn, kmax,
                                          mu.ctypes.data_as(C.POINTER(C.c_ubyte)) if mu is not None else None,
                                          _d(out), C.byref(tot), _d(post), _d(nc), _d(sp), _d(sl), _d(nl))
    assert rc == 0
    return out, Result(tot.value, post, nc, sp, sl, nl)


def given_configs(L: Locus, configs: np.ndarray):
    """postcal.cpp:400-714 (-b/-d/-e): configs = int16[num_configs][num_groups] of global SNP indices, negative = none.
    Returns (rc, Result): rc 0 ok, 2 = the reference's "This did not work as expected" exit, 3 = out-of-range entry."""
    n, sig, z, d, smap = L.flat()
    cfg = np.ascontiguousarray(configs, dtype=np.int16)
    nn, ng = cfg.shape
    post = np.zeros(L.N); nc = np.zeros(L.S); sp = np.zeros(L.U); sl = np.zeros(L.U); nl = np.zeros(L.U)
    tot = C.c_double(); ne = _u64()
    rc = lib().oracle_given_configs(L.S, _i(n), _d(sig), _d(z), C.c_double(L.K), _d(d), L.U, _i(smap),
                                    C.c_double(L.gamma), C.c_double(L.p), cfg.ctypes.data_as(C.POINTER(C.c_int16)),
                                    C.c_int64(nn), int(ng), C.byref(tot), _d(post), _d(nc), _d(sp), _d(sl), _d(nl),
                                    C.byref(ne))
    return rc, Result(tot.value, post, nc, sp, sl, nl, int(ne.value))


def sss(L: Locus, c: int, max_iter: int = 1000, trace_cap: int = 1000) -> Result:
    n, sig, z, d, smap = L.flat()
    post = np.zeros(L.N); nc = np.zeros(L.S); sp = np.zeros(L.U); sl = np.zeros(L.U); nl = np.zeros(L.U)
    tot = C.c_double(); nit = C.c_int()
    trace = np.full((trace_cap, 4), -1, dtype=np.int64)
    rc = lib().oracle_sss(L.S, _i(n), _d(sig), _d(z), C.c_double(L.K), _d(d), L.U, _i(smap), int(c),
                          C.c_double(L.gamma), C.c_double(L.p), int(max_iter), C.byref(tot), _d(post), _d(nc),
                          _d(sp), _d(sl), _d(nl), C.byref(nit), trace.ctypes.data_as(C.POINTER(C.c_int64)), trace_cap)
    assert rc == 0
    used = trace[trace[:, 0] >= 0]
    return Result(tot.value, post, nc, sp, sl, nl, extra={"n_iter": nit.value, "trace": used})


def unrank(rank, U, c):
    out = np.zeros(max(c, 1), dtype=np.int32)
    k = lib().oracle_unrank(_u64(rank), int(U), int(c), _i(out))
    return out[:k].tolist()


def walk(U, steps):
    out = np.zeros(U, dtype=np.int32)
    lib().oracle_walk(int(U), _u64(steps), _i(out))
    return np.nonzero(out)[0].tolist()


def expansions(snp_map, locs):
    smap = np.ascontiguousarray(snp_map, dtype=np.int32)
    locs = np.ascontiguousarray(locs, dtype=np.int32)
    nc = len(locs)
    cap = 3 ** nc
    out = np.zeros((cap, max(nc, 1)), dtype=np.int32)
    k = lib().oracle_expansions(smap.shape[1], _i(smap), _i(locs), nc, _i(out), cap)
    return out[:k, :nc]


def f_block(sigma, z, d, Cset):
    Cset = np.ascontiguousarray(Cset, dtype=np.int32)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    return lib().oracle_f_block(sigma.shape[0], _d(sigma), _d(z), C.c_double(d), _i(Cset), len(Cset))


def ll_dense(B, Sp, d, Cset):
    Cset = np.ascontiguousarray(Cset, dtype=np.int32)
    B = np.ascontiguousarray(B, dtype=np.float64)
    Sp = np.ascontiguousarray(Sp, dtype=np.float64)
    return lib().oracle_ll_dense(B.shape[0], _d(B), _d(Sp), C.c_double(d), _i(Cset), len(Cset))


def ref_dir():
    return os.path.join(_HERE, "_ref")


def parse_raw_dump(path):
    """Parse <out>_raw.txt written by oracle/_ref/pipsort_ref_dump (oracle/ref_dump.cpp)."""
    total = K = None
    S = N = U = 0
    post = nc = sp = sl = nl = None
    with open(path) as f:
        for line in f:
            t = line.split()
            if t[0] == "total":
                total = float(t[1])
            elif t[0] == "K":
                K = float(t[1])
            elif t[0] == "dims":
                S, N, U = int(t[1]), int(t[2]), int(t[3])
                post = np.zeros(N); nc = np.zeros(S); sp = np.zeros(U); sl = np.zeros(U); nl = np.zeros(U)
            elif t[0] == "noCausal":
                nc[int(t[1])] = float(t[2])
            elif t[0] == "postValues":
                post[int(t[1])] = float(t[2])
            elif t[0] == "shared":
                g = int(t[1]); sp[g] = float(t[2]); sl[g] = float(t[3]); nl[g] = float(t[4])
    return Result(total, post, nc, sp, sl, nl, extra={"K": K})
