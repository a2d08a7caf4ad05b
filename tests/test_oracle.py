"""Pin the CPU oracle (oracle/oracle.cpp) against the reference.

* tests/golden/*.json  : 17-digit dumps of the UNMODIFIED reference (oracle/_ref, make_golden.py)
* tests/golden/example/expected_* : the reference's own shipped golden files (6 printed digits)
* tests/golden/small_example/mscaviar_results_log.txt line 1 (exhaustive c=3 p=0.75)
"""
import itertools
import os

import numpy as np
import pytest

from conftest import GOLDEN, args_to_params, given_config_matrix, golden, has_golden, oracle_locus
from oracle import oracle as O

# log-likelihoods within 1e-10 relative, PIPs within 1e-8 absolute (BASELINE.json north_star)
RTOL_LL = 1e-10
ATOL_PIP = 1e-8


def check_against_golden(r, g, rtol=RTOL_LL):
    assert r.total == pytest.approx(g["total"], rel=rtol)
    for name in ["post", "noCausal", "sharedPips", "sharedLL", "notSharedLL"]:
        got, want = getattr(r, name), np.array(g[name])
        assert np.array_equal(got == 0, want == 0), name          # same "empty" pattern
        np.testing.assert_allclose(got, want, rtol=rtol, atol=0, err_msg=name)
    ref = O.Result(g["total"], np.array(g["post"]), np.array(g["noCausal"]), np.array(g["sharedPips"]),
                   np.array(g["sharedLL"]), np.array(g["notSharedLL"]))
    with np.errstate(over="ignore"):
        np.testing.assert_allclose(r.pips(), ref.pips(), atol=ATOL_PIP, rtol=0)
        np.testing.assert_allclose(r.shared_pips(), ref.shared_pips(), atol=ATOL_PIP, rtol=0)
        np.testing.assert_allclose(r.no_causal(), ref.no_causal(), atol=ATOL_PIP, rtol=0)


EXH = ["small_c1_p075", "small_c2_p025", "small_c2_p075", "small_c3_p075", "small_c3_p0", "small_c3_g005_t1_s3",
       "example_c1_p025", "example_c2_p025"]


@pytest.mark.parametrize("name", EXH)
def test_exhaustive_matches_reference_dump(name):
    if not has_golden(name):
        pytest.skip("golden not generated")
    g = golden(name)
    prm = args_to_params(g["args"])
    L = oracle_locus(g["dataset"], p=prm["p"], gamma=prm["gamma"], s=prm["s"], t=prm["t"])
    assert L.K == pytest.approx(g["K"], rel=1e-12)
    r = O.exhaustive(L, prm["c"])
    check_against_golden(r, g)


GIVEN = ["small_given_72x5", "small_given_mixed_p025", "example_given_mixed"]


@pytest.mark.parametrize("name", GIVEN)
def test_given_configs_match_reference_dump(name):
    """-b/-d/-e (postcal.cpp:400-714) against 17-digit dumps of the reference on the same int16 matrices."""
    g = golden(name)
    prm = args_to_params(g["args"])
    L = oracle_locus(g["dataset"], p=prm["p"], gamma=prm["gamma"], s=prm["s"], t=prm["t"])
    cfg, _ = given_config_matrix(prm)
    rc, r = O.given_configs(L, cfg)
    assert rc == 0 and r.n_eval == prm["d"]
    check_against_golden(r, g)


def test_given_configs_rejects_what_the_reference_rejects():
    L = oracle_locus("small_example")
    ok = np.array([[0, 5, -1]], dtype=np.int16)
    assert O.given_configs(L, ok)[0] == 0
    assert O.given_configs(L, np.array([[5, 0, -1]], dtype=np.int16))[0] == 2      # out of order: "did not work"
    assert O.given_configs(L, np.array([[3, 3, -1]], dtype=np.int16))[0] == 2      # duplicate entry
    assert O.given_configs(L, np.array([[0, 11, -1]], dtype=np.int16))[0] == 3     # >= N


def test_counts():
    # 76 / 268 configurations on small_example, 216,817 on example
    Ls = oracle_locus("small_example")
    assert O.exhaustive(Ls, 2).n_eval == 76
    assert O.exhaustive(Ls, 3).n_eval == 268
    Le = oracle_locus("example", p=0.25)
    assert O.total_union_subsets(Le.U, 2) == 24977
    assert O.exhaustive(Le, 2).n_eval == 216817


def test_small_example_log_file():
    with open(os.path.join(GOLDEN, "small_example", "mscaviar_results_log.txt")) as f:
        want = f.readline().strip()
    r = O.exhaustive(oracle_locus("small_example"), 3)
    assert "%g" % np.exp(r.total) == want == "3.99066e-18"


def fmt6(x):
    """default ostream << double = %g with 6 significant digits."""
    return "%g" % x


def test_example_expected_files():
    """The six files the reference ships for tests/example, from the oracle's numbers."""
    L = oracle_locus("example", p=0.25)
    r = O.exhaustive(L, 2)
    d = os.path.join(GOLDEN, "example")
    pips, off = r.pips(), 0
    for s in range(2):
        lines = ["SNP_ID\tProb_in_pCausalSet"] + [f"{nm}\t{fmt6(pips[off + i])}" for i, nm in enumerate(L.names[s])]
        with open(os.path.join(d, f"expected_study{s}_post.txt")) as f:
            assert f.read().splitlines() == lines
        sel = [nm for i, nm in enumerate(L.names[s]) if pips[off + i] > 0.05]        # postcal.cpp:1158-1163
        with open(os.path.join(d, f"expected_study{s}_set.txt")) as f:
            assert f.read().splitlines() == sel
        off += len(L.names[s])
    with open(os.path.join(d, "expected_nocausal.txt")) as f:
        assert f.read().splitlines() == [fmt6(v) for v in r.no_causal()]
    with np.errstate(over="ignore"):
        sp = r.shared_pips()
    lines = ["SNP_ID\tshared_pip\tshared_ll\tnotshared_ll"] + [
        f"{nm}\t{fmt6(sp[g])}\t{fmt6(r.sharedLL[g])}\t{fmt6(r.notSharedLL[g])}" for g, nm in enumerate(L.union_names)]
    with open(os.path.join(d, "expected_shared_pips.txt")) as f:
        assert f.read().splitlines() == lines


@pytest.mark.parametrize("name", ["small_sss_c3_p075", "small_sss_c2_p025", "example_sss_c2_p025"])
def test_sss_matches_reference_dump(name):
    if not has_golden(name):
        pytest.skip("golden not generated")
    g = golden(name)
    prm = args_to_params(g["args"])
    L = oracle_locus(g["dataset"], p=prm["p"], gamma=prm["gamma"], s=prm["s"], t=prm["t"])
    r = O.sss(L, prm["c"])
    check_against_golden(r, g)


def test_sss_small_trace():
    # SURVEY 8a S3 [probe]: iter 0 nbd 10 (10 new) -> {7}; iter 1 nbd 19 (10 new) -> {}; iter 2 nothing new
    r = O.sss(oracle_locus("small_example"), 3)
    tr = r.extra["trace"]
    assert tr[:, 1].tolist() == [10, 19, 10] and tr[:, 2].tolist() == [10, 10, 0]
    assert "%g" % np.exp(r.total) == "1.88456e-18"


# ---- enumeration ------------------------------------------------------------------------------
@pytest.mark.parametrize("U,c", [(5, 3), (10, 2), (7, 7), (12, 4)])
def test_unrank_equals_walk_and_itertools(U, c):
    seq = [list(cmb) for j in range(c + 1) for cmb in itertools.combinations(range(U), j)]
    assert O.total_union_subsets(U, c) == len(seq)
    for r, want in enumerate(seq):
        assert O.unrank(r, U, c) == want
        assert O.walk(U, r) == want


def test_unrank_large():
    U, c = 6000, 5
    tot = O.total_union_subsets(U, c)
    assert O.unrank(tot - 1, U, c) == [5995, 5996, 5997, 5998, 5999]
    assert O.unrank(1 + U, U, c) == [0, 1]
    import math
    assert O.lib().oracle_ncr(6000, 5) == math.comb(6000, 5)


def test_expansion_order():
    # slots: causal j outer, study inner; bit t of bmask -> slot t; checkOR (postcal.cpp:928-958)
    smap = np.array([[0, 1, -1, 2], [0, -1, 1, 2]], dtype=np.int32)
    e = O.expansions(smap, [0, 3])
    # lowest chosen union index is the fastest digit, digit order s0 < s1 < both
    assert e.tolist() == [[1, 1], [2, 1], [3, 1], [1, 2], [2, 2], [3, 2], [1, 3], [2, 3], [3, 3]]
    assert O.expansions(smap, [1, 2]).tolist() == [[1, 2]]
    assert O.expansions(smap, [0, 1, 2]).tolist() == [[1, 1, 2], [2, 1, 2], [3, 1, 2]]


# ---- likelihood -------------------------------------------------------------------------------
def test_closed_form_equals_dense_woodbury():
    """f_block (O(k^3) closed form) == lowrank_likelihood's dense Woodbury form from B and S'."""
    ld, z = O.read_ld(os.path.join(GOLDEN, "small_example", "eur_small_test.ld")), None
    _, z = O.read_z(os.path.join(GOLDEN, "small_example", "eur_small_test_final.zscore"))
    sig, ze, K, add, B, Sp = O.preprocess(ld, z)
    assert add == pytest.approx(0.01)                         # singular LD -> one 0.01 step
    np.testing.assert_allclose(ze, z, rtol=1e-12)             # B^T S' == z
    for k in range(1, 5):
        for C in itertools.combinations(range(5), k):
            a = -K / 2 + O.f_block(sig, ze, 5.72, C)
            b = O.ll_dense(B, Sp, 5.72, C)
            assert a == pytest.approx(b, rel=1e-12)


def test_sharded_ranks_merge():
    """Splitting the rank space and merging in log space reproduces the whole run."""
    L = oracle_locus("small_example")
    whole = O.exhaustive(L, 3)
    tot = O.total_union_subsets(L.U, 3)
    cut = tot // 3
    a, b = O.exhaustive(L, 3, 0, cut), O.exhaustive(L, 3, cut, tot)
    assert a.n_eval + b.n_eval == whole.n_eval
    merged = np.logaddexp(a.total, b.total)
    assert merged == pytest.approx(whole.total, rel=1e-13)
    t, n = O.exhaustive_omp(L, 3, 0, tot, 3)
    assert n == whole.n_eval and t == pytest.approx(whole.total, rel=1e-13)
