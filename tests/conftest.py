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
    prm = dict(c=3, p=0.75, gamma=0.01, t=0.52, s=5.2, q=0)
    key = {"-c": "c", "-p": "p", "-g": "gamma", "-t": "t", "-s": "s", "-q": "q"}
    for k, v in zip(args[::2], args[1::2]):
        prm[key[k]] = int(v) if k in ("-c", "-q") else float(v)
    return prm


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
