"""Parity of the CUDA engine (through the C-ABI) with the CPU oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): log-likelihoods 1e-10 relative, PIPs / shared PIPs / no-causal
probabilities 1e-8 absolute, enumeration and credible-set membership bit-exact.
"""
import itertools
import os

import numpy as np
import pytest

from conftest import (GOLDEN, args_to_params, assert_results_match, engine_for, golden, has_golden, oracle_locus,
                      synth_as_oracle_locus)

pytestmark = pytest.mark.gpu

EXH = ["small_c1_p075", "small_c2_p025", "small_c2_p075", "small_c3_p075", "small_c3_p0", "small_c3_g005_t1_s3",
       "example_c1_p025", "example_c2_p025"]


@pytest.mark.parametrize("keep_order", [False, True])
@pytest.mark.parametrize("name", EXH)
def test_exhaustive_matches_reference_dump(name, keep_order):
    if not has_golden(name):
        pytest.skip("golden not generated")
    g = golden(name)
    prm = args_to_params(g["args"])
    L = oracle_locus(g["dataset"], p=prm["p"], gamma=prm["gamma"], s=prm["s"], t=prm["t"])
    with engine_for(L, prm["c"], keep_order=keep_order) as e:
        r = e.compute_total_likelihood(prm["c"])
    assert_results_match(r, g)


def test_generic_only_flag_matches():
    """PIPSORT_GENERIC_ONLY routes every size class through the warp-per-configuration kernel."""
    from oracle import oracle as O
    L = oracle_locus("small_example")
    want = O.exhaustive(L, 3)
    with engine_for(L, 3, generic_only=True) as e:
        r = e.compute_total_likelihood(3)
    assert r.n_configs == want.n_eval
    assert_results_match(r, want)


def test_counts_and_oracle_small():
    from oracle import oracle as O
    L = oracle_locus("small_example")
    with engine_for(L, 3) as e:
        for c, n in [(0, 1), (1, 12), (2, 76), (3, 268)]:
            r = e.compute_total_likelihood(c)
            assert r.n_configs == O.exhaustive(L, c).n_eval
            if c >= 2:
                assert r.n_configs == n
            assert_results_match(r, O.exhaustive(L, c))


def fmt6(x):
    return "%g" % x


def test_example_expected_files():
    """The six files the reference ships for tests/example (run_example.sh), from the engine's numbers."""
    L = oracle_locus("example", p=0.25)
    with engine_for(L, 2) as e:
        r = e.compute_total_likelihood(2)
    assert r.n_configs == 216817
    d = os.path.join(GOLDEN, "example")
    pips, off = r.pips(), 0
    for s in range(2):
        lines = ["SNP_ID\tProb_in_pCausalSet"] + [f"{nm}\t{fmt6(pips[off + i])}" for i, nm in enumerate(L.names[s])]
        with open(os.path.join(d, f"expected_study{s}_post.txt")) as f:
            assert f.read().splitlines() == lines
        sel = [nm for i, nm in enumerate(L.names[s]) if pips[off + i] > 0.05]        # postcal.cpp:1158-1163
        with open(os.path.join(d, f"expected_study{s}_set.txt")) as f:
            assert f.read().splitlines() == sel                                        # credible set: bit-exact
        off += len(L.names[s])
    with open(os.path.join(d, "expected_nocausal.txt")) as f:
        assert f.read().splitlines() == [fmt6(v) for v in r.no_causal()]
    sp = r.shared_pips()
    lines = ["SNP_ID\tshared_pip\tshared_ll\tnotshared_ll"] + [
        f"{nm}\t{fmt6(sp[g])}\t{fmt6(r.sharedLL[g])}\t{fmt6(r.notSharedLL[g])}" for g, nm in enumerate(L.union_names)]
    with open(os.path.join(d, "expected_shared_pips.txt")) as f:
        assert f.read().splitlines() == lines


def test_partial_rank_ranges_keep_order():
    """With PIPSORT_KEEP_ORDER the rank space is the reference's: partial ranges equal the oracle's."""
    from oracle import oracle as O
    L = oracle_locus("small_example")
    tot = O.total_union_subsets(L.U, 3)
    with engine_for(L, 3, keep_order=True) as e:
        assert e.total_ranks(3) == tot
        for lo, hi in [(0, 1), (0, 11), (5, 60), (56, 57), (30, tot), (tot - 1, tot), (7, 7)]:
            e.reset()
            e.run_exhaustive(3, lo, hi)
            r = e.read()
            want = O.exhaustive(L, 3, lo, hi)
            assert r.n_configs == want.n_eval
            assert_results_match(r, want)


def test_shards_merge_to_whole():
    from oracle import oracle as O
    L = oracle_locus("small_example")
    whole = O.exhaustive(L, 3)
    for parts in (2, 3, 8):
        engines = [engine_for(L, 3) for _ in range(parts)]
        b = engines[0].shard_ranks(3, parts)
        assert b[0] == 0 and b[-1] == engines[0].total_ranks(3) and all(x <= y for x, y in zip(b, b[1:]))
        for i, e in enumerate(engines):
            e.run_exhaustive(3, b[i], b[i + 1])
        for e in engines[1:]:
            engines[0].merge_from(e)
        r = engines[0].read()
        assert r.n_configs == whole.n_eval
        assert_results_match(r, whole)
        for e in engines:
            e.close()


def test_enumeration_bit_exact():
    """rank -> union subset and expansion -> per-study states, against the oracle's walk (nextBinary) and mask loop."""
    from oracle import oracle as O
    L = oracle_locus("small_example")
    with engine_for(L, 3) as e:
        tot = e.total_ranks(3)
        seq = [list(cmb) for j in range(4) for cmb in itertools.combinations(range(L.U), j)]
        assert tot == len(seq)
        for rank in list(range(0, 40)) + [56, 57, 100, tot - 1]:
            idx, st, ne = e.enumerate(3, rank, 0)
            assert idx == seq[rank] == O.walk(L.U, rank)
            ex = O.expansions(L.snp_map, idx) if idx else np.zeros((1, 0), dtype=np.int32)
            assert ne == len(ex)
            for x in range(ne):
                _, st, _ = e.enumerate(3, rank, x)
                assert st == ex[x].tolist()
    # a large rank space: U = 6000, c = 5 unranks without a walk
    smap = np.stack([np.arange(6000), np.arange(6000)]).astype(np.int32)
    import pipsort_b200 as P
    n = np.array([6000, 6000], dtype=np.int32)
    # tiny LD (identity) -- only the map matters for enumeration
    with P.Engine([1, 1], [np.eye(1), np.eye(1)], [np.zeros(1), np.zeros(1)], [1.0, 1.0], 0.0,
                  np.zeros((2, 6000), dtype=np.int32) - 1, max_causal=5) as e2:
        tot = e2.total_ranks(5)
        assert tot == O.total_union_subsets(6000, 5)
        for rank in [0, 1, 6000, 6001, 12345678901, tot - 1]:
            idx, _, _ = e2.enumerate(5, rank, 0)
            assert idx == O.unrank(rank, 6000, 5)


def _random_union_configs(rng, U, n, kmax):
    idx = np.full((n, kmax), -1, dtype=np.int32)
    for i in range(n):
        k = int(rng.integers(0, kmax + 1))
        idx[i, :k] = np.sort(rng.choice(U, k, replace=False))
    return idx


@pytest.mark.parametrize("dataset,kmax", [("small_example", 3), ("small_example", 5), ("example", 4)])
def test_score_union_configs_matches_oracle(dataset, kmax):
    """expand_and_compute_lkl for a batch: max-|l| outputs and accumulators, with make_updates on/off."""
    from oracle import oracle as O
    L = oracle_locus(dataset)
    rng = np.random.default_rng(7)
    idx = _random_union_configs(rng, L.U, 64, kmax)
    idx[0, :] = -1                                   # the null configuration
    upd = (rng.uniform(size=64) < 0.7).astype(np.uint8)
    upd[0] = 1
    want_l, want = O.score_union_configs(L, idx, upd)
    with engine_for(L, kmax) as e:
        got_l = e.score_union_configs(idx, upd)
        r = e.read()
        np.testing.assert_allclose(got_l, want_l, rtol=1e-10, atol=0)
        assert_results_match(r, want)
        # a second batch accumulates on top of the first (the SSS keeps one PostCal across iterations)
        idx2 = _random_union_configs(rng, L.U, 33, kmax)
        want_l2, want2 = O.score_union_configs(L, idx2, None, want)
        got_l2 = e.score_union_configs(idx2)
        np.testing.assert_allclose(got_l2, want_l2, rtol=1e-10, atol=0)
        assert_results_match(e.read(), want2)


def test_generic_and_register_paths_agree():
    """All union subsets of size <= 3 scored through the generic (warp per configuration) kernel equal the
    exhaustive driver's result (two independently written CUDA paths)."""
    L = oracle_locus("small_example")
    seq = [list(cmb) for j in range(4) for cmb in itertools.combinations(range(L.U), j)]
    idx = np.full((len(seq), 3), -1, dtype=np.int32)
    for i, s in enumerate(seq):
        idx[i, :len(s)] = s
    with engine_for(L, 3) as e:
        a = e.compute_total_likelihood(3)
        e.reset()
        e.score_union_configs(idx)
        b = e.read()
    assert a.n_configs == b.n_configs
    assert_results_match(a, b, rtol=1e-12)


@pytest.mark.parametrize("n,overlap,c,p", [(12, 0.5, 3, 0.75), (40, 0.8, 3, 0.25), (60, 1.0, 2, 0.75), (30, 0.0, 3, 0.5),
                                           (33, 0.8, 3, 0.0)])
def test_synthetic_matches_oracle(n, overlap, c, p):
    """Mixed SNP types (shared / study-specific), both SNP orders, against the oracle."""
    from oracle import oracle as O
    from pipsort_b200 import synth
    SL = synth.make_locus(n, overlap=overlap, seed=1234 + n, sharing_param=p)
    L = synth_as_oracle_locus(SL)
    want = O.exhaustive(L, c)
    assert want.n_eval == synth.count_configs(SL.snp_map, c)
    for keep in (False, True):
        with engine_for(SL, c, keep_order=keep) as e:
            r = e.compute_total_likelihood(c)
        assert r.n_configs == want.n_eval
        assert_results_match(r, want)


def test_full_size_properties():
    """BASELINE configs at full size (150 SNPs/study c=3: 12,197,751 configurations; 300/study c=2) through
    size-independent properties: configuration count, shard additivity, sum rules of the accumulators."""
    from pipsort_b200 import synth
    for n, c in [(150, 3), (300, 2)]:
        SL = synth.make_locus(n, overlap=0.8)
        with engine_for(SL, c) as e:
            r = e.compute_total_likelihood(c)
            assert r.n_configs == synth.count_configs(SL.snp_map, c)
            # every configuration has between 1 and c causal union SNPs, except the null one:
            #   sum_g (X1+X2+X3)[g] = sum_conf j(conf) w(conf)  in [total - null, c (total - null)]
            pips = r.pips()
            assert np.all(pips >= 0) and np.all(pips <= 1 + 1e-12)
            nc = r.no_causal()
            assert np.all(nc >= 0) and np.all(nc <= 1)
            sp = r.shared_pips()
            # shared PIP <= PIP in either study
            smap = SL.snp_map
            for g in range(SL.U):
                if smap[0, g] >= 0 and smap[1, g] >= 0:
                    assert sp[g] <= pips[smap[0, g]] + 1e-12 and sp[g] <= pips[n + smap[1, g]] + 1e-12
                else:
                    assert sp[g] == 0
            # shards add up
            b = e.shard_ranks(c, 4)
            e.reset()
            for i in range(4):
                e.run_exhaustive(c, b[i], b[i + 1])
            r2 = e.read()
            assert r2.n_configs == r.n_configs
            assert_results_match(r2, r, rtol=1e-12)


def test_graph_replay_reproduces_the_pass():
    """reset + exhaustive + finalize recorded into a CUDA graph (pipsort_graph_begin/end) and replayed: identical results,
    and the accumulators do not pile up across replays (the recorded reset is part of the graph)."""
    from oracle import oracle as O
    L = oracle_locus("small_example")
    want = O.exhaustive(L, 3)
    with engine_for(L, 3) as e:
        first = e.compute_total_likelihood(3)          # allocates the scratch buffers outside the capture
        e.graph_begin()
        e.reset(); e.run_exhaustive(3); e.finalize()
        gid = e.graph_end()
        for _ in range(3):
            e.graph_launch(gid)
        r = e.read()
    assert r.n_configs == first.n_configs == 268
    assert_results_match(r, want)
    assert r.total == pytest.approx(first.total, rel=1e-13)


@pytest.mark.parametrize("p", [0.25, 0.75])
def test_baseline_config_a_300_snps_c2(p):
    """BASELINE.json configs[2] at full size: synthetic 300-SNP/study locus, 80 % overlap (U = 360), c = 2 exhaustive
    (352,501 configurations), p = 0.25 and 0.75 -- every accumulator against the oracle, as one run and as the 1 / 2 / 4 / 8
    rank-range shards the multi-GPU path hands out (accumulated on one engine: the stores are additive)."""
    from oracle import oracle as O
    from pipsort_b200 import synth
    SL = synth.make_locus(300, overlap=0.8, sharing_param=p)
    assert SL.U == 360 and synth.count_configs(SL.snp_map, 2) == 352501
    want = O.exhaustive(synth_as_oracle_locus(SL), 2)
    with engine_for(SL, 2) as e:
        r = e.compute_total_likelihood(2)
        assert r.n_configs == want.n_eval == 352501
        assert_results_match(r, want)
        for parts in (2, 4, 8):
            b = e.shard_ranks(2, parts)
            e.reset()
            for i in range(parts):
                e.run_exhaustive(2, b[i], b[i + 1])
            rs = e.read()
            assert rs.n_configs == 352501
            assert_results_match(rs, want)


def test_baseline_config_b_150_snps_c3_sampled_against_oracle():
    """BASELINE.json configs[3] (150 SNPs/study, c = 3, 12,197,751 configurations): the oracle needs a few seconds for the
    whole run, so compare everything."""
    from oracle import oracle as O
    from pipsort_b200 import synth
    SL = synth.make_locus(150, overlap=0.8)
    want_total, want_n = O.exhaustive_omp(synth_as_oracle_locus(SL), 3, 0, O.total_union_subsets(SL.U, 3), 0)
    with engine_for(SL, 3) as e:
        r = e.compute_total_likelihood(3)
    assert r.n_configs == want_n == 12197751
    assert r.total == pytest.approx(want_total, rel=1e-10)


def test_one_call_posterior_matches_engine_path():
    """pipsort_posterior_exhaustive (create + run + read + destroy in one C-ABI call) == the step-by-step path."""
    import pipsort_b200 as P
    g = golden("example_c2_p025")
    L = oracle_locus("example", p=0.25)
    r = P.posterior_exhaustive(L.n_snps, L.sigma, L.z, L.d, L.K, L.snp_map, 2, gamma=L.gamma, sharing_param=L.p)
    assert r.n_configs == 216817
    assert_results_match(r, g)


def test_batch_of_loci_pipelined():
    """pipsort_posterior_exhaustive_batch: different loci in one call (three engines in flight on three streams); every
    locus must come back exactly as the one-at-a-time path computes it."""
    import pipsort_b200 as P
    from oracle import oracle as O
    from pipsort_b200 import synth
    loci, want = [], []
    for i, (n, ov, p) in enumerate([(12, 0.5, 0.75), (40, 0.8, 0.25), (30, 0.0, 0.5), (25, 1.0, 0.75), (60, 0.7, 0.3), (33, 0.8, 0.0),
                                    (12, 0.5, 0.75)]):
        SL = synth.make_locus(n, overlap=ov, seed=100 + i, sharing_param=p)
        loci.append(dict(num_snps=SL.num_snps, sigma=np.concatenate([s.ravel() for s in SL.sigma]), z=np.concatenate(SL.z),
                         d=SL.d, K=SL.K, snp_map=SL.snp_map, gamma=SL.gamma, sharing_param=p))
        want.append(O.exhaustive(synth_as_oracle_locus(SL), 3))
    got = P.posterior_exhaustive_batch(loci, 3)
    assert len(got) == len(want)
    for r, w in zip(got, want):
        assert r.n_configs == w.n_eval
        assert_results_match(r, w)
    assert P.posterior_exhaustive_batch([], 3) == []


def test_edge_cases_empty_and_ragged():
    """Empty union, a study without SNPs, a single SNP, c larger than U: the counts and sums of the oracle."""
    import pipsort_b200 as P
    from oracle import oracle as O
    from pipsort_b200 import synth
    # no SNP at all: only the null configuration exists (postcal.cpp:793-822)
    with P.Engine([0, 0], np.zeros(0), np.zeros(0), [5.72, 5.72], 3.0, np.zeros((2, 0), dtype=np.int32), max_causal=3) as e:
        r = e.compute_total_likelihood(3)
        assert r.n_configs == 1 and r.total == pytest.approx(-3.0 / 2 - 1.0, rel=1e-14)
        assert r.noCausal.tolist() == [r.total, r.total]
    # ragged: study 1 has no SNP, study 0 has four; c = 3 > number of shared SNPs (0)
    SL = synth.make_locus(4, overlap=0.0, seed=9)
    smap = np.array([[0, 1, 2, 3], [-1, -1, -1, -1]], dtype=np.int32)
    L = O.Locus(n_snps=np.array([4, 0], dtype=np.int32), sigma=[SL.sigma[0], np.zeros((0, 0))], z=[SL.z[0], np.zeros(0)], K=1.25,
                d=SL.d, snp_map=smap, gamma=0.01, p=0.75)
    want = O.exhaustive(L, 3)
    with engine_for(L, 3) as e:
        r = e.compute_total_likelihood(3)
    assert r.n_configs == want.n_eval == 1 + 4 + 6 + 4
    assert_results_match(r, want)
    # one shared SNP, c = 3 > U = 1
    L1 = O.Locus(n_snps=np.array([1, 1], dtype=np.int32), sigma=[np.ones((1, 1)), np.ones((1, 1))], z=[np.array([2.5]), np.array([-1.0])],
                 K=7.25, d=np.array([5.72, 5.72]), snp_map=np.zeros((2, 1), dtype=np.int32), gamma=0.01, p=0.75)
    want = O.exhaustive(L1, 3)
    with engine_for(L1, 3) as e:
        r = e.compute_total_likelihood(3)
    assert r.n_configs == want.n_eval == 4
    assert_results_match(r, want)


def test_out_of_range_locus_is_refused():
    """z-scores so large that exp(f) leaves every representable range: PIPSORT_E_RANGE at create, not garbage later."""
    import pipsort_b200 as P
    with pytest.raises(P.PipsortError) as ei:
        P.Engine([1, 1], [np.ones((1, 1)), np.ones((1, 1))], [np.array([1e6]), np.array([1.0])], [257.0, 5.72], 1.0,
                 np.zeros((2, 1), dtype=np.int32), max_causal=3)
    assert ei.value.code == 4


@pytest.mark.parametrize("dataset,c,p", [("small_example", 3, 0.75), ("example", 2, 0.25)])
def test_finalize_reset_and_fetch(dataset, c, p):
    """pipsort_finalize_reset: bins -> results AND an empty store in one launch (few bins: lane-per-bin path; tests/example:
    ~80 bins per accumulator, the scanning path); pipsort_fetch_results returns what the last finalize produced.  Passes
    repeated without pipsort_reset must not pile up."""
    from oracle import oracle as O
    L = oracle_locus(dataset, p=p)
    want = O.exhaustive(L, c) if dataset == "small_example" else golden("example_c2_p025")
    with engine_for(L, c) as e:
        for rep in range(3):
            e.run_exhaustive(c)
            e.finalize(reset=True)
            r = e.fetch()
            assert r.n_configs == (268 if dataset == "small_example" else 216817)
            assert_results_match(r, want)
        empty = e.read()                               # the store is empty now
        assert empty.n_configs == 0 and empty.total == 0.0
        assert not empty.postValues.any() and not empty.sharedLL.any() and not empty.noCausal.any()
        assert_results_match(e.fetch(), empty)         # ... and fetch returns the last finalize (that of read())
