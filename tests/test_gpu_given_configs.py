"""Explicit-configuration path (-b/-d/-e; PostCal::computeTotalLikelihoodGivenConfigs, postcal.cpp:400-714) on the GPU,
through the C-ABI (pipsort_score_given_configs), against 17-digit dumps of the reference and the CPU oracle."""
import numpy as np
import pytest

from conftest import (args_to_params, assert_results_match, engine_for, given_config_matrix, golden, oracle_locus,
                      synth_as_oracle_locus)

pytestmark = pytest.mark.gpu

GIVEN = ["small_given_72x5", "small_given_mixed_p025", "example_given_mixed"]


@pytest.mark.parametrize("keep_order", [False, True])
@pytest.mark.parametrize("name", GIVEN)
def test_given_configs_match_reference_dump(name, keep_order):
    g = golden(name)
    prm = args_to_params(g["args"])
    L = oracle_locus(g["dataset"], p=prm["p"], gamma=prm["gamma"], s=prm["s"], t=prm["t"])
    cfg, _ = given_config_matrix(prm)
    with engine_for(L, 3, keep_order=keep_order) as e:
        r = e.compute_total_likelihood_given_configs(cfg)
    assert r.n_configs == prm["d"]
    assert_results_match(r, g)


def random_rows(rng, N, rows, groups, kmax):
    M = np.full((rows, groups), -1, dtype=np.int16)
    for r in range(rows):
        k = int(rng.integers(0, kmax + 1))
        M[r, np.sort(rng.choice(groups, size=k, replace=False))] = np.sort(rng.choice(N, size=k, replace=False))
    return M


def test_given_configs_synthetic_matches_oracle():
    """40+40 SNPs, mixed SNP types, 5000 rows of up to 8 causal SNPs (more than any exhaustive run reaches)."""
    from oracle import oracle as O
    from pipsort_b200 import synth
    L = synth.make_locus(40, overlap=0.6, seed=11)
    cfg = random_rows(np.random.default_rng(5), L.N, 5000, 9, 8)
    rc, want = O.given_configs(synth_as_oracle_locus(L), cfg)
    assert rc == 0
    with engine_for(L, 3) as e:
        r = e.compute_total_likelihood_given_configs(cfg)
        assert r.n_configs == 5000
        assert_results_match(r, want)
        # accumulation across calls == one call on the concatenation (PostCal's arrays persist, postcal.h:129-160)
        e.reset()
        e.score_given_configs(cfg[:1234])
        e.score_given_configs(cfg[1234:])
        assert_results_match(e.read(), want)


def test_given_configs_equal_exhaustive_enumeration():
    """Feeding the explicit path every configuration the exhaustive walk visits reproduces the exhaustive result
    (two independently written CUDA paths + the reference's enumeration, bit-exact through pipsort_enumerate)."""
    L = oracle_locus("small_example")
    off = [0, int(L.n_snps[0])]
    with engine_for(L, 3, keep_order=True) as e:
        want = e.compute_total_likelihood(3)
        rows = []
        for rank in range(e.total_ranks(3)):
            idx, st, ne = e.enumerate(3, rank, 0)
            for x in range(ne):
                idx, st, _ = e.enumerate(3, rank, x)
                ent = []
                for g, t in zip(idx, st):
                    if t in (1, 3):
                        ent.append(off[0] + int(L.snp_map[0][g]))
                    if t in (2, 3):
                        ent.append(off[1] + int(L.snp_map[1][g]))
                ent = sorted(ent)
                rows.append(ent + [-1] * (6 - len(ent)))
        cfg = np.array(rows, dtype=np.int16)
        assert cfg.shape[0] == want.n_configs == 268
        got = e.compute_total_likelihood_given_configs(cfg)
    assert_results_match(got, want)


def test_given_configs_rejects_bad_rows():
    import pipsort_b200 as P
    L = oracle_locus("small_example")
    for bad in ([[5, 0, -1]], [[3, 3, -1]], [[0, 11, -1]]):
        with engine_for(L, 3) as e:
            with pytest.raises(P.PipsortError) as ei:
                e.score_given_configs(np.array(bad, dtype=np.int16))
            assert ei.value.code == 6 and "did not work as expected" in str(ei.value)
