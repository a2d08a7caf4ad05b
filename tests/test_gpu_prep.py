"""Model's pre-processing (model.h:171-264: PSD shift by LU determinant, eigen-decomposition, |Omega|, B, S') on the GPU
(pipsort_preprocess_study / PIPSORT_RAW_LD) against the oracle's restatement, which is pinned on the reference's golden
files through exactly this step (a_s, K and the effective LD shape every number the reference prints)."""
import numpy as np
import pytest

from conftest import DATASETS, assert_results_match, dataset_paths, engine_for, golden, oracle_locus, args_to_params

pytestmark = pytest.mark.gpu


def raw_inputs(dataset):
    from oracle import oracle as O
    ld_files, z_files, mp, n = dataset_paths(dataset)
    lds = [O.read_ld(f) for f in ld_files]
    zs = [O.read_z(f)[1] for f in z_files]
    return lds, zs


@pytest.mark.parametrize("dataset", ["small_example", "example"])
def test_preprocess_study_matches_oracle(dataset):
    import pipsort_b200 as P
    from oracle import oracle as O
    lds, zs = raw_inputs(dataset)
    L = oracle_locus(dataset)
    Ksum = 0.0
    for s, (ld, z) in enumerate(zip(lds, zs)):
        want_sig, _, want_K, want_add, _, _ = O.preprocess(ld, z)
        got_sig, info = P.preprocess_study(ld, z)
        assert info["add_diag"] == want_add                       # same number of 0.01 steps, accumulated the same way
        assert info["K"] == pytest.approx(want_K, rel=1e-11)
        np.testing.assert_allclose(got_sig, want_sig, rtol=0, atol=1e-12)
        Ksum += info["K"]
    assert Ksum == pytest.approx(L.K, rel=1e-11)
    # small_example: study 0 has two identical LD rows (singular -> +0.01), study 1 needs no shift
    if dataset == "small_example":
        assert P.preprocess_study(lds[0], zs[0])[1]["add_diag"] == 0.01
        assert P.preprocess_study(lds[1], zs[1])[1]["add_diag"] == 0.0


@pytest.mark.parametrize("name", ["small_c3_p075", "example_c2_p025"])
def test_raw_ld_engine_matches_reference_dump(name):
    """Raw LD in, everything on the device: PSD shift + eigen + exhaustive run == the reference's 17-digit dump."""
    import pipsort_b200 as P
    g = golden(name)
    prm = args_to_params(g["args"])
    lds, zs = raw_inputs(g["dataset"])
    L = oracle_locus(g["dataset"], p=prm["p"], gamma=prm["gamma"], s=prm["s"], t=prm["t"])
    with P.Engine(L.n_snps, lds, zs, L.d, 0.0, L.snp_map, gamma=L.gamma, sharing_param=L.p, max_causal=prm["c"],
                  raw_ld=True) as e:
        r = e.compute_total_likelihood(prm["c"])
        assert sum(e.prep_info(s)["K"] for s in range(2)) == pytest.approx(g["K"], rel=1e-11)
    assert_results_match(r, g)


def test_negative_eigenvalues_are_flipped():
    """An indefinite 'LD' with an EVEN number of negative eigenvalues has a positive determinant, so the PSD loop
    adds nothing and model.h:227 takes |Omega|: effective LD = Q |Omega| Q^T, K = sum (q.z)^2 / |Omega|."""
    import pipsort_b200 as P
    from oracle import oracle as O
    rng = np.random.default_rng(7)
    n = 24
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    w = np.concatenate([[-0.3, -0.05], rng.uniform(0.2, 2.0, n - 2)])
    ld = (Q * w) @ Q.T
    ld = (ld + ld.T) / 2
    z = rng.standard_normal(n) * 2
    got_sig, info = P.preprocess_study(ld, z)
    assert info["add_diag"] == 0.0 and info["n_negative"] == 2
    want_sig = (Q * np.abs(w)) @ Q.T
    np.testing.assert_allclose(got_sig, want_sig, atol=1e-12)
    qz = Q.T @ z
    assert info["K"] == pytest.approx(float(np.sum(qz * qz / np.abs(w))), rel=1e-11)
    o_sig, _, o_K, o_add, _, _ = O.preprocess(ld, z)
    assert o_add == 0.0
    np.testing.assert_allclose(got_sig, o_sig, atol=1e-11)
    assert info["K"] == pytest.approx(o_K, rel=1e-10)


def test_underflowing_determinant_keeps_shifting():
    """util.cpp:204-221: a determinant that underflows to 0 is 'not positive'.  500 SNPs with pivots around 0.2 have a
    determinant far below 1e-324, so the loop must keep adding 0.01 although the matrix is positive definite."""
    import pipsort_b200 as P
    from oracle import oracle as O
    from pipsort_b200 import synth
    L = synth.make_locus(500, overlap=0.8, seed=5)
    ld, z = L.sigma[0], L.z[0]
    assert np.linalg.slogdet(ld)[1] < -745                          # below the smallest positive double
    got_sig, info = P.preprocess_study(ld, z)
    _, _, o_K, o_add, _, _ = O.preprocess(ld, z)
    assert info["add_diag"] == o_add and info["add_diag"] > 0
    assert info["psd_iterations"] == round(o_add / 0.01) + 1
    assert info["K"] == pytest.approx(o_K, rel=1e-10)
    np.testing.assert_allclose(got_sig, ld + o_add * np.eye(500), atol=1e-12)


def test_slightly_asymmetric_ld_follows_the_reference():
    """An LD file that is not exactly symmetric (values printed to 6 digits, estimated separately above and below the
    diagonal): the reference LU-factorises the matrix AS READ (util.cpp:204-215) and eigen-decomposes its LOWER triangle
    (gsl_eigen_symmv, util.cpp:242), so B^T B is the symmetric completion of the lower triangle (+ the shift).  The
    device path must do the same: same shift, K, and an exactly symmetric effective LD equal to the oracle's."""
    import pipsort_b200 as P
    from oracle import oracle as O
    from pipsort_b200 import synth
    L = synth.make_locus(40, overlap=0.8, seed=11)
    rng = np.random.default_rng(3)
    for s in range(2):
        ld = L.sigma[s].copy()
        ld += np.triu(rng.uniform(-2e-4, 2e-4, ld.shape), 1)        # upper triangle perturbed, lower untouched
        if s == 1:
            ld[7] = ld[3]                                           # two identical rows: singular as read -> the loop must shift
            ld[:, 7] = ld[:, 3]
            ld += np.triu(rng.uniform(-1e-4, 1e-4, ld.shape), 1)
        z = L.z[s]
        want_sig, _, want_K, want_add, _, _ = O.preprocess(ld, z)
        got_sig, info = P.preprocess_study(ld, z)
        assert info["add_diag"] == want_add
        assert np.array_equal(got_sig, got_sig.T)
        np.testing.assert_allclose(got_sig, want_sig, rtol=0, atol=1e-11)
        assert info["K"] == pytest.approx(want_K, rel=1e-9)
