"""Seeded synthetic two-ancestry loci (SURVEY.md section 8d "Synthetic inputs").

Per study: genotype-like matrix G (n_ref x N_s) with AR(1) structure inside LD blocks of 10-25 SNPs,
LD = corr(G) (positive definite: n_ref = 2 N_s), z = LD @ lambda + eps with eps ~ N(0, LD) and one
shared + one study-specific causal effect.  The snp_map places a random `overlap` fraction of each
study's SNPs in both studies (rows: rsid, idx0 or -1, idx1 or -1; each study's SNPs appear in
increasing study-index order, as model.h:134-144 requires).

What is produced is what crosses the engine boundary (effective LD, z, K, d, snp_map); `write_files`
also emits the reference's input file formats so the unmodified reference can run the same locus.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class SynthLocus:
    num_snps: np.ndarray          # int32[2]
    sigma: list                   # per study float64[n,n]
    z: list                       # per study float64[n]
    K: float
    d: np.ndarray                 # float64[2]
    snp_map: np.ndarray           # int32[2,U]
    gamma: float
    sharing_param: float
    sample_sizes: tuple
    names: list = field(default_factory=list)        # per study SNP names
    union_names: list = field(default_factory=list)

    @property
    def U(self):
        return self.snp_map.shape[1]

    @property
    def N(self):
        return int(self.num_snps.sum())

    def n_types(self):
        sh = int(((self.snp_map[0] >= 0) & (self.snp_map[1] >= 0)).sum())
        o0 = int(((self.snp_map[0] >= 0) & (self.snp_map[1] < 0)).sum())
        o1 = int(((self.snp_map[0] < 0) & (self.snp_map[1] >= 0)).sum())
        return sh, o0, o1


def count_configs(snp_map, c):
    """Number of expanded configurations = sum over union subsets of size <= c of 3^#shared,
    by elementary symmetric polynomials of the per-SNP state counts."""
    smap = np.asarray(snp_map)
    w = np.where((smap[0] >= 0) & (smap[1] >= 0), 3, np.where((smap[0] >= 0) | (smap[1] >= 0), 1, 0))
    e = [1] + [0] * c
    for x in w.tolist():
        for m in range(c, 0, -1):
            e[m] += x * e[m - 1]
    return int(sum(e))


def flops_per_config_total(snp_map, c):
    """Algorithmic FP64 flops of the whole exhaustive run, SURVEY.md section 8(d):
    per study block phi(k) = k^3/3 + 2.5 k^2 + 31/6 k + 2 (k > 0), per configuration
    phi(k0)+phi(k1) + 5 + [1 + 1{k0=0} + 1{k1=0} + k0 + k1 + 2 j].  Exact, by class counting."""
    from math import comb
    smap = np.asarray(snp_map)
    sh = int(((smap[0] >= 0) & (smap[1] >= 0)).sum())
    o0 = int(((smap[0] >= 0) & (smap[1] < 0)).sum())
    o1 = int(((smap[0] < 0) & (smap[1] >= 0)).sum())

    def phi(k):
        return 0.0 if k == 0 else k ** 3 / 3.0 + 2.5 * k * k + 31.0 / 6.0 * k + 2.0

    total = 0.0
    nconf = 0
    for a in range(c + 1):                 # shared SNPs chosen
        for b0 in range(c + 1 - a):        # study-0-only SNPs chosen
            for b1 in range(c + 1 - a - b0):
                j = a + b0 + b1
                nsub = comb(sh, a) * comb(o0, b0) * comb(o1, b1)
                if nsub == 0 or j == 0:
                    continue
                # states of the a shared SNPs: x in study 0 only, y in study 1 only, rest both
                for x in range(a + 1):
                    for y in range(a + 1 - x):
                        both = a - x - y
                        mult = comb(a, x) * comb(a - x, y)
                        k0, k1 = x + both + b0, y + both + b1
                        f = phi(k0) + phi(k1) + 5 + (1 + (k0 == 0) + (k1 == 0) + k0 + k1 + 2 * j)
                        total += nsub * mult * f
                        nconf += nsub * mult
    return total, nconf


def flops_of_union_subset(a, b0, b1):
    """(algorithmic flops, expanded configurations) of ONE union subset with a shared, b0 study-0-only and b1 study-1-only
    SNPs -- the per-configuration count of flops_per_config_total."""
    from math import comb

    def phi(k):
        return 0.0 if k == 0 else k ** 3 / 3.0 + 2.5 * k * k + 31.0 / 6.0 * k + 2.0

    j = a + b0 + b1
    if j == 0:
        return 8.0, 1          # the null configuration: prior, shift, exp + three accumulator adds
    total, n = 0.0, 0
    for x in range(a + 1):
        for y in range(a + 1 - x):
            both = a - x - y
            mult = comb(a, x) * comb(a - x, y)
            k0, k1 = x + both + b0, y + both + b1
            total += mult * (phi(k0) + phi(k1) + 5 + (1 + (k0 == 0) + (k1 == 0) + k0 + k1 + 2 * j))
            n += mult
    return total, n


def _block_ld(rng, n, n_ref):
    G = np.empty((n_ref, n))
    i = 0
    while i < n:
        b = int(rng.integers(10, 26))
        b = min(b, n - i)
        rho = 0.9
        e = rng.standard_normal((n_ref, b))
        x = np.empty((n_ref, b))
        x[:, 0] = e[:, 0]
        for t in range(1, b):
            x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho * rho) * e[:, t]
        G[:, i:i + b] = x
        i += b
    G -= G.mean(0)
    G /= G.std(0)
    return (G.T @ G) / n_ref


def make_locus(n_per_study=150, overlap=0.8, seed=20261018, sample_sizes=(100000, 20000), gamma=0.01,
               sharing_param=0.75, s_squared=5.2, t_squared=0.52, round_ld=False) -> SynthLocus:
    rng = np.random.default_rng(seed)
    n = [int(n_per_study), int(n_per_study)]
    n_sh = int(round(overlap * n_per_study))
    sig, zs = [], []
    # which SNPs of each study are shared (sorted so the map keeps each study's order)
    sh_idx = [np.sort(rng.choice(n[s], n_sh, replace=False)) for s in range(2)]
    for s in range(2):
        ld = _block_ld(rng, n[s], 2 * n[s])
        if round_ld:
            ld = np.array([[float("%g" % v) for v in row] for row in ld])
            ld = (ld + ld.T) / 2
        sig.append(ld)
    # causal effects: one shared SNP (same union SNP in both studies) + one study-specific each
    lam = [np.zeros(n[0]), np.zeros(n[1])]
    if n_sh > 0:
        t = int(rng.integers(0, n_sh))
        lam[0][sh_idx[0][t]] = rng.uniform(5, 8)
        lam[1][sh_idx[1][t]] = rng.uniform(5, 8)
    for s in range(2):
        lam[s][int(rng.integers(0, n[s]))] += rng.uniform(5, 8)
    K = 0.0
    for s in range(2):
        Lc = np.linalg.cholesky(sig[s])
        z = sig[s] @ lam[s] + Lc @ rng.standard_normal(n[s])
        if round_ld:
            z = np.array([float("%g" % v) for v in z])
        zs.append(z)
        y = np.linalg.solve(Lc, z)
        K += float(y @ y)
    # union list: merge the two studies' SNPs by a common "position" so shared SNPs interleave
    rows = []  # (position key, idx0, idx1)
    shared0 = {int(i): t for t, i in enumerate(sh_idx[0])}
    pos0 = np.sort(rng.uniform(0, 1, n[0]))
    pos1_free = np.sort(rng.uniform(0, 1, n[1]))
    for i in range(n[0]):
        if i in shared0:
            rows.append((pos0[i], i, int(sh_idx[1][shared0[i]])))
        else:
            rows.append((pos0[i], i, -1))
    sh1 = set(int(i) for i in sh_idx[1])
    for i in range(n[1]):
        if i not in sh1:
            rows.append((pos1_free[i], -1, i))
    # keep each study's SNPs in increasing study-index order (model.h:134-144): sort by (idx0 order) is not
    # enough for study 1, so place rows by study-0 order first and insert study-1-only rows where their
    # index falls relative to the shared study-1 indices.
    a_rows = [r for r in rows if r[1] >= 0]
    b_rows = sorted([r for r in rows if r[1] < 0], key=lambda r: r[2])
    out, bi = [], 0
    for r in a_rows:
        if r[2] >= 0:
            while bi < len(b_rows) and b_rows[bi][2] < r[2]:
                out.append(b_rows[bi]); bi += 1
        out.append(r)
    out.extend(b_rows[bi:])
    smap = np.array([[r[1] for r in out], [r[2] for r in out]], dtype=np.int32)
    mn = int(min(sample_sizes))
    d = np.array([s_squared * (float(x) / mn) + t_squared for x in sample_sizes])
    names = [[f"s{s}_rs{i}" for i in range(n[s])] for s in range(2)]
    union_names = [f"rs{g}" for g in range(smap.shape[1])]
    return SynthLocus(np.array(n, dtype=np.int32), sig, zs, K, d, smap, gamma, sharing_param, tuple(sample_sizes),
                      names, union_names)


def write_files(L: SynthLocus, outdir: str):
    """Reference input formats: LD = whitespace separated doubles, z = 'name z' per line,
    snp_map = 'rsid,idx0,idx1', plus the -l / -z list files.  Returns (ldfiles, zfiles, map, n-string)."""
    os.makedirs(outdir, exist_ok=True)
    ldl, zl = [], []
    for s in range(2):
        lp, zp = os.path.join(outdir, f"s{s}.ld"), os.path.join(outdir, f"s{s}.zscores")
        with open(lp, "w") as f:
            for row in L.sigma[s]:
                f.write(" ".join("%.17g" % v for v in row) + "\n")
        with open(zp, "w") as f:
            for nm, v in zip(L.names[s], L.z[s]):
                f.write(f"{nm} {v:.17g}\n")
        ldl.append(lp); zl.append(zp)
    with open(os.path.join(outdir, "ldfiles.txt"), "w") as f:
        f.write("\n".join(ldl) + "\n")
    with open(os.path.join(outdir, "zfiles.txt"), "w") as f:
        f.write("\n".join(zl) + "\n")
    mp = os.path.join(outdir, "snp_map")
    with open(mp, "w") as f:
        for g, nm in enumerate(L.union_names):
            f.write(f"{nm},{L.snp_map[0, g]},{L.snp_map[1, g]}\n")
    return os.path.join(outdir, "ldfiles.txt"), os.path.join(outdir, "zfiles.txt"), mp, ",".join(str(x) for x in L.sample_sizes)
