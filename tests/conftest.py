import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def has_golden(name):
    return os.path.exists(os.path.join(GOLDEN, name + ".json"))


DATASETS = {
    "small_example": dict(
        ld=["eur_small_test.ld", "afr_small_test.ld"],
        z=["eur_small_test_final.zscore", "afr_small_test_final.zscore"],
        map="eur_afr_small_test_snp_map", n=[7000, 7000]),
    "example": dict(ld=["s1.ld", "s2.ld"], z=["s1_final.zscores", "s2_final.zscores"], map="snp_map",
                    n=[334324, 6771]),
}


def dataset_paths(name):
    ds = DATASETS[name]
    d = os.path.join(GOLDEN, name)
    return ([os.path.join(d, f) for f in ds["ld"]], [os.path.join(d, f) for f in ds["z"]],
            os.path.join(d, ds["map"]), ds["n"])


def args_to_params(args):
    """['-c','2','-p','0.25', ...] -> dict with the reference's defaults (pipsort.cpp:69-77)."""
    prm = dict(c=3, p=0.75, gamma=0.01, t=0.52, s=5.2, q=0, b=None, d=0, e=0)
    key = {"-c": "c", "-p": "p", "-g": "gamma", "-t": "t", "-s": "s", "-q": "q", "-b": "b", "-d": "d", "-e": "e"}
    for k, v in zip(args[::2], args[1::2]):
        prm[key[k]] = v if k == "-b" else (int(v) if k in ("-c", "-q", "-d", "-e") else float(v))
    return prm


def given_config_matrix(prm):
    """The int16 matrix a golden case was run on ('@rel/path' is relative to tests/golden)."""
    import numpy as np
    path = prm["b"]
    path = os.path.join(GOLDEN, path[1:]) if path.startswith("@") else path
    return np.fromfile(path, dtype=np.int16).reshape(prm["d"], prm["e"]), path


_locus_cache = {}


def oracle_locus(dataset, p=0.75, gamma=0.01, s=5.2, t=0.52):
    """Locus built by the ORACLE's restatement of the host pre-processing (cached per dataset)."""
    from oracle import oracle as O
    import copy
    if dataset not in _locus_cache:
        ld, z, mp, n = dataset_paths(dataset)
        _locus_cache[dataset] = O.load_locus(ld, z, mp, n)
    L = copy.copy(_locus_cache[dataset])
    L.p, L.gamma = p, gamma
    L.d = O.d_per_study(DATASETS[dataset]["n"], s, t)
    return L


@pytest.fixture(scope="session")
def small_locus():
    return oracle_locus("small_example")


@pytest.fixture(scope="session")
def example_locus():
    return oracle_locus("example", p=0.25)


# ---- GPU side ------------------------------------------------------------------------------------------
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def engine_for(L, max_causal=3, **kw):
    """pipsort_b200.Engine from an oracle Locus / SynthLocus (what crosses the C-ABI boundary)."""
    import pipsort_b200 as P
    n = getattr(L, "n_snps", None)
    if n is None:
        n = L.num_snps
    p = getattr(L, "p", None)
    if p is None:
        p = L.sharing_param
    return P.Engine(n, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=p, max_causal=max_causal, **kw)


def synth_as_oracle_locus(SL):
    from oracle import oracle as O
    return O.Locus(n_snps=SL.num_snps, sigma=SL.sigma, z=SL.z, K=SL.K, d=SL.d, snp_map=SL.snp_map, gamma=SL.gamma,
                   p=SL.sharing_param)


# log-likelihoods within 1e-10 relative, PIPs within 1e-8 absolute (BASELINE.json north_star)
RTOL_LL = 1e-10
ATOL_PIP = 1e-8


def assert_results_match(got, want, rtol=RTOL_LL, atol_pip=ATOL_PIP):
    """got: pipsort_b200.Results (or oracle Result); want: oracle Result / golden dict."""
    import numpy as np
    if isinstance(want, dict):
        from oracle import oracle as O
        want = O.Result(want["total"], np.array(want["post"]), np.array(want["noCausal"]), np.array(want["sharedPips"]),
                        np.array(want["sharedLL"]), np.array(want["notSharedLL"]))
    gp = getattr(got, "postValues", None)
    if gp is None:
        gp = got.post
    if not hasattr(want, "post"):
        want.post = want.postValues
    assert got.total == pytest.approx(want.total, rel=rtol)
    pairs = [("post", gp, want.post), ("noCausal", got.noCausal, want.noCausal),
             ("sharedPips", got.sharedPips, want.sharedPips), ("sharedLL", got.sharedLL, want.sharedLL),
             ("notSharedLL", got.notSharedLL, want.notSharedLL)]
    for name, a, b in pairs:
        a, b = np.asarray(a), np.asarray(b)
        assert np.array_equal(a == 0, b == 0), f"{name}: empty pattern differs"
        np.testing.assert_allclose(a, b, rtol=rtol, atol=0, err_msg=name)
    with np.errstate(over="ignore"):
        def sx(x, t):
            return np.where(np.asarray(x) == 0, 0.0, np.exp(np.asarray(x) - t))
        np.testing.assert_allclose(sx(gp, got.total), sx(want.post, want.total), atol=atol_pip, rtol=0)
        np.testing.assert_allclose(sx(got.sharedPips, got.total), sx(want.sharedPips, want.total), atol=atol_pip, rtol=0)
        np.testing.assert_allclose(sx(got.noCausal, got.total), sx(want.noCausal, want.total), atol=atol_pip, rtol=0)


_synth_cache = {}


def synth_locus(n, overlap=0.8, seed=20261018, sharing_param=0.75):
    """pipsort_b200.synth.make_locus, cached per session (the 1500- and 5000-SNP loci take seconds to generate)."""
    from pipsort_b200 import synth
    key = (n, overlap, seed, sharing_param)
    if key not in _synth_cache:
        _synth_cache[key] = synth.make_locus(n, overlap=overlap, seed=seed, sharing_param=sharing_param)
    return _synth_cache[key]


class exh_plan_env:
    """Force the work decomposition of the exhaustive launch for the duration of a with-block: PIPSORT_EXH_CHUNK = target
    cost of a chunk in warp-steps (read per launch by the planner, csrc/exhaustive.cuh; None = the planner's own choice)."""

    def __init__(self, chunk=None):
        self.want = {"PIPSORT_EXH_CHUNK": chunk}
        self.old = {}

    def __enter__(self):
        for k, v in self.want.items():
            self.old[k] = os.environ.get(k)
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        return self

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
