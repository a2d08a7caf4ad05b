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


def _reference_psd_shift(ld):
    """makeSigmaPositiveSemiDefinite (util.cpp:195-226) with SciPy's LU: the running product of the pivots in index order."""
    import scipy.linalg as sl
    add, it = 0.0, 0
    n = ld.shape[0]
    while True:
        lu, piv = sl.lu_factor(ld + add * np.eye(n))
        d = -1.0 if (np.count_nonzero(piv != np.arange(n)) % 2) else 1.0
        for v in np.diag(lu):
            d *= v
        it += 1
        if d > 0:
            return add, it
        add += 0.01


@pytest.mark.parametrize("n,force", [(600, True), (1100, False)])
def test_cholesky_certificates_keep_the_reference_shift(n, force):
    """The PSD loop with most LU factorizations replaced by Cholesky certificates (prep.cuh) must stop at exactly the shift
    the reference's loop stops at (here: SciPy's partial-pivoting LU + the running pivot product, which underflows for
    hundreds of SNPs), and K / the effective LD must be those of the eigen path.  n = 600 forces the certificate path on a
    size the oracle's own pre-processing can check; n = 1100 takes it by default."""
    import subprocess, sys, json, os
    from pipsort_b200 import synth
    L = synth.make_locus(n, overlap=0.8, seed=5)
    ld, z = L.sigma[0], L.z[0]
    want_add, want_it = _reference_psd_shift(ld)
    assert want_it > 2
    # the threshold is read once per process: run the forced case in a child
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import pipsort_b200 as P; from pipsort_b200 import synth;"
            "L = synth.make_locus(%d, overlap=0.8, seed=5); s, i = P.preprocess_study(L.sigma[0], L.z[0]);"
            "print(json.dumps(dict(info=i, dev=float(np.abs(s - (L.sigma[0] + i['add_diag'] * np.eye(%d))).max()), sym=bool((s == s.T).all()))))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), n, n))
    env = dict(os.environ, PIPSORT_PREP_CHOL_MIN="64" if force else "1024", PIPSORT_TRACE_PREP="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert "Cholesky certificates" in p.stderr and " 0 Cholesky certificates" not in p.stderr, p.stderr
    assert out["info"]["add_diag"] == want_add and out["info"]["psd_iterations"] == want_it
    assert out["dev"] <= 1e-12 and out["sym"]
    K = float(z @ np.linalg.solve(ld + want_add * np.eye(n), z))
    assert out["info"]["K"] == pytest.approx(K, rel=1e-10)
    if force:
        from oracle import oracle as O
        _, _, o_K, o_add, _, _ = O.preprocess(ld, z)
        assert o_add == want_add and out["info"]["K"] == pytest.approx(o_K, rel=1e-10)
