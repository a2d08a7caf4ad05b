#!/usr/bin/env python
"""Generate the full-precision golden vectors under tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/pipsort_ref_dump (reference sources compiled where they lie, see oracle/Makefile and
oracle/ref_dump.cpp) on the reference's own example inputs (copied verbatim as data into
tests/golden/{example,small_example}) and stores the 17-digit log-space arrays as JSON.

Can only run in the build container (needs /root/reference to have produced oracle/_ref); the JSON
files it writes are committed so that nothing at test time needs the reference.

Exhaustive cases are run with OMP_THREAD_LIMIT=1: the reference updates noCausal / sharedPips /
sharedLL / notSharedLL without synchronisation inside its 64-thread loop (postcal.cpp:998,1012-1016)
and loses updates run to run (observed: notSharedLL differing in the 6th digit between two runs of
tests/small_example).  One thread gives the race-free value the code intends.  The SSS path holds
`omp critical` around every update (sss_postcal.cpp:629-667) and is run with all cores.

usage: python tests/golden/make_golden.py [case ...]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

DUMP = os.path.join(ROOT, "oracle", "_ref", "pipsort_ref_dump")

SMALL = dict(dir="small_example", map="eur_afr_small_test_snp_map", n="7000,7000")
EX = dict(dir="example", map="snp_map", n="334324,6771")

CASES = {
    # name: (dataset, extra args, serial?)
    "small_c1_p075": (SMALL, ["-c", "1", "-p", "0.75"], True),
    "small_c2_p025": (SMALL, ["-c", "2", "-p", "0.25"], True),
    "small_c2_p075": (SMALL, ["-c", "2", "-p", "0.75"], True),
    "small_c3_p075": (SMALL, ["-c", "3", "-p", "0.75"], True),          # run_test.sh:2 (defaults)
    "small_c3_p0": (SMALL, ["-c", "3", "-p", "0"], True),              # p == 0 skips the sharing term
    "small_c3_g005_t1_s3": (SMALL, ["-c", "3", "-p", "0.5", "-g", "0.05", "-t", "1.0", "-s", "3.0"], True),
    "small_sss_c3_p075": (SMALL, ["-c", "3", "-p", "0.75", "-q", "1"], False),   # run_test.sh:1
    "small_sss_c2_p025": (SMALL, ["-c", "2", "-p", "0.25", "-q", "1"], False),
    "example_c1_p025": (EX, ["-c", "1", "-p", "0.25"], True),
    "example_c2_p025": (EX, ["-c", "2", "-p", "0.25"], True),           # run_example.sh:1
    "example_sss_c2_p025": (EX, ["-c", "2", "-p", "0.25", "-q", "1"], False),
    # explicit configurations (-b/-d/-e, postcal.cpp:400-714); the OpenMP loop there updates noCausal / sharedPips /
    # sharedLL / notSharedLL without synchronisation too (:481-483, :650, :664-669) -> one thread
    "small_given_72x5": (SMALL, ["-b", "@test_optional_configs/all_configs_int16", "-d", "72", "-e", "5"], True),   # tests/test_optional_configs/run_test.sh:1
    "small_given_mixed_p025": (SMALL, ["-b", "@test_optional_configs/small_mixed_int16", "-d", "96", "-e", "6", "-p", "0.25"], True),
    "example_given_mixed": (EX, ["-b", "@test_optional_configs/example_mixed_int16", "-d", "240", "-e", "7", "-p", "0.25"], True),
}


def make_config_matrix(path, n_snps, rows, groups, seed):
    """A seeded int16 matrix in the layout utils/construct_configs_all_studies.py:104-158 writes: every row lists
    global SNP indices (offset_s + i) in increasing order with -1 for unused groups; includes all-(-1) rows, rows
    with several causal SNPs per study and rows that hit the same union SNP in both studies."""
    import numpy as np
    rng = np.random.default_rng(seed)
    N = int(sum(n_snps))
    M = np.full((rows, groups), -1, dtype=np.int16)
    for r in range(rows):
        k = int(rng.integers(0, groups + 1)) if r % 11 else 0
        pick = np.sort(rng.choice(N, size=k, replace=False))
        cols = np.sort(rng.choice(groups, size=k, replace=False))
        M[r, cols] = pick
    M.tofile(path)
    return M


def run_case(name):
    ds, extra, serial = CASES[name]
    d = os.path.join(HERE, ds["dir"])
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    if serial:
        env["OMP_THREAD_LIMIT"] = "1"
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "o")
        xargs = [os.path.join(HERE, a[1:]) if a.startswith("@") else a for a in extra]
        cmd = [DUMP, "-l", "ldfiles.txt", "-z", "zfiles.txt", "-m", ds["map"], "-n", ds["n"], "-o", out] + xargs
        t = time.time()
        p = subprocess.run(cmd, cwd=d, env=env, capture_output=True, text=True)
        dt = time.time() - t
        assert p.returncode == 0, p.stderr[-2000:]
        r = O.parse_raw_dump(out + "_raw.txt")
        files = {}
        for suf in ["study0_post", "study1_post", "study0_set", "study1_set", "nocausal", "shared_pips", "log"]:
            with open(f"{out}_{suf}.txt") as f:
                files[suf] = f.read()
        flags = [ln for ln in p.stdout.splitlines() if "hit " in ln]
    js = dict(name=name, dataset=ds["dir"], args=extra, sample_sizes=ds["n"], serial=serial, seconds=round(dt, 2),
              total=r.total, K=r.extra["K"], post=r.post.tolist(), noCausal=r.noCausal.tolist(),
              sharedPips=r.sharedPips.tolist(), sharedLL=r.sharedLL.tolist(), notSharedLL=r.notSharedLL.tolist(),
              stdout_flags=flags, files=files)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(js, f, indent=0)
    print(f"{name}: total={r.total!r}  {dt:.1f}s  {flags}")


if __name__ == "__main__":
    assert os.path.exists(DUMP), "build oracle/_ref first: make -C oracle ref"
    oc = os.path.join(HERE, "test_optional_configs")
    if not os.path.exists(os.path.join(oc, "small_mixed_int16")):
        make_config_matrix(os.path.join(oc, "small_mixed_int16"), [5, 6], 96, 6, 1)
    if not os.path.exists(os.path.join(oc, "example_mixed_int16")):
        make_config_matrix(os.path.join(oc, "example_mixed_int16"), [222, 219], 240, 7, 2)
    for nm in (sys.argv[1:] or CASES.keys()):
        run_case(nm)
