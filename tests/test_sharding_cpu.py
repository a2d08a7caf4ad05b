"""Host side of the multi-GPU path on CPU: world_size-2 `gloo` process groups (no GPU needed).

What runs for real here: the shard planner of the C-ABI (pipsort_shard_ranks_for_map, pure host arithmetic) and the
torch.distributed plumbing of pipsort_b200/distributed.py (shard -> run -> ONE all-reduce(sum) -> read; neighbourhood
slices + all-gather for the shotgun search).  The CUDA engine is replaced by a stand-in that fills the same kind of
additive accumulator vector from the CPU oracle (test infrastructure), so the merged result can be compared with the
oracle's whole-run result.  The same functions are driven with the real engine + NCCL by bench.py --gpus N.
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, assert_results_match, oracle_locus

import pipsort_b200 as P
from pipsort_b200 import distributed as D
from pipsort_b200 import synth


class OracleBackedEngine:
    """Stand-in with the Engine methods distributed.py uses.  Accumulators are LINEAR sums exp(l) (small_example's
    log-likelihoods are about -40, so plain doubles suffice): additive across shards exactly like the binned store."""

    def __init__(self, L, c):
        from oracle import oracle as O
        self.O, self.L, self.max_causal = O, L, c
        self.n = 2 + L.N + L.S + 3 * L.U            # total, count, post, noCausal, sharedPips, sharedLL, notSharedLL
        self.acc = torch.zeros(self.n, dtype=torch.float64)

    def shard_ranks(self, c, parts):                # the real planner; snp_map order = the oracle's rank order
        return P.engine.shard_ranks_for_map(self.L.snp_map, c, parts, keep_order=True)

    def reset(self):
        self.acc.zero_()

    def _add(self, r, n_eval):
        lin = lambda x: np.where(np.asarray(x) == 0, 0.0, np.exp(np.asarray(x, dtype=np.float64)))  # noqa: E731
        v = np.concatenate([[float(lin(r.total)), float(n_eval)], lin(r.post), lin(r.noCausal), lin(r.sharedPips),
                            lin(r.sharedLL), lin(r.notSharedLL)])
        self.acc += torch.from_numpy(v)

    def run_exhaustive(self, c, lo, hi):
        if hi > lo:
            r = self.O.exhaustive(self.L, c, lo, hi)
            self._add(r, r.n_eval)

    def score_union_configs(self, idx, make_updates=None):
        out, r = self.O.score_union_configs(self.L, idx, make_updates)
        self._add(r, 0)
        return out

    def accumulator_tensor(self):
        return self.acc

    def read(self):
        L = self.L
        v = self.acc.numpy()
        lg = lambda x: np.where(x > 0, np.log(np.where(x > 0, x, 1.0)), 0.0)  # noqa: E731
        o = 2
        post = lg(v[o:o + L.N]); o += L.N
        nc = lg(v[o:o + L.S]); o += L.S
        sp = lg(v[o:o + L.U]); o += L.U
        sl = lg(v[o:o + L.U]); o += L.U
        nl = lg(v[o:o + L.U])
        return P.Results(float(lg(v[0:1])[0]), post, nc, sp, sl, nl, int(round(v[1])))


def _worker(rank, world, initfile, outdir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        L = oracle_locus("small_example")
        e = OracleBackedEngine(L, 3)
        res = D.compute_total_likelihood_sharded(e, 3)
        want = O.exhaustive(L, 3)
        assert res.n_configs == want.n_eval == 268
        assert_results_match(res, want, rtol=1e-12)
        # a caller-supplied split (here: deliberately lopsided) must give the same sums
        res2 = D.compute_total_likelihood_sharded(e, 3, bounds=[0, 7, O.total_union_subsets(L.U, 3)])
        assert_results_match(res2, want, rtol=1e-12)

        # shotgun-search neighbourhood: slices + gather of the scores, accumulators merged on read
        idx = np.array([[-1, -1, -1], [7, -1, -1], [0, 7, -1], [1, 4, 9], [2, 3, -1], [5, -1, -1], [0, 1, 2]], dtype=np.int32)
        e.reset()
        got = D.score_union_configs_sharded(e, idx)
        want_l, want_acc = O.score_union_configs(L, idx)
        np.testing.assert_allclose(got, want_l, rtol=1e-12)
        merged = D.read_sharded(e)
        assert_results_match(merged, want_acc, rtol=1e-12)
        part = e.read()                     # the store itself is still this rank's partial sum
        assert part.total < merged.total
        open(os.path.join(outdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_run_matches_whole_run():
    with tempfile.TemporaryDirectory() as tmp:
        initfile = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(2, initfile, tmp), nprocs=2, join=True)
        assert os.path.exists(os.path.join(tmp, "ok0")) and os.path.exists(os.path.join(tmp, "ok1"))


@pytest.mark.parametrize("parts", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("keep_order", [False, True])
def test_shard_planner_partitions_and_balances(parts, keep_order):
    """Bounds are monotone, cover [0,total) and split the kernel's work (cost-weighted warp-steps) evenly."""
    L = synth.make_locus(40, overlap=0.5, seed=3)
    c = 3
    U = L.U
    b = P.engine.shard_ranks_for_map(L.snp_map, c, parts, keep_order=keep_order)
    from math import comb
    total = sum(comb(U, j) for j in range(c + 1))
    assert b[0] == 0 and b[-1] == total and all(x <= y for x, y in zip(b, b[1:]))
    # the register kernel's unit of work is a warp-step (a, b, 32-wide tile of x); a step of a tile whose x's are in one
    # study only costs 0.4-0.7 of a generic one (ExhCostModel, exh_plan.h), so a shard may hold up to ~2.5x its share of subsets
    sizes = [b[i + 1] - b[i] for i in range(parts)]
    assert max(sizes) <= 2.5 * total / parts + 64
    assert min(sizes) >= 0.3 * total / parts - 64
    if keep_order:                                  # rank order == snp_map order: the shards partition the configurations
        from oracle import oracle as O
        from conftest import synth_as_oracle_locus
        OL = synth_as_oracle_locus(L)
        work = [O.exhaustive_omp(OL, c, b[i], b[i + 1], 2)[1] for i in range(parts)]
        assert sum(work) == synth.count_configs(L.snp_map, c)


def test_slice_bounds():
    assert D.slice_bounds(10, 3) == [0, 3, 6, 10]
    assert D.slice_bounds(0, 4) == [0, 0, 0, 0, 0]
    assert D.slice_bounds(5, 8)[-1] == 5
