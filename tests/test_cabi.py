"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/pipsort_b200.h declares, and the product fails loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def declared_symbols():
    with open(os.path.join(ROOT, "include", "pipsort_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pipsort_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import pipsort_b200 as P
    lib = P.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pipsort_b200.h but not exported"
    assert "sm_100a" in P.version()


def test_no_cpu_fallback():
    if have_gpu():
        pytest.skip("GPU present")
    import pipsort_b200 as P
    with pytest.raises(P.PipsortError) as ei:
        P.Engine([1, 1], [np.eye(1), np.eye(1)], [np.zeros(1), np.zeros(1)], [1.0, 1.0], 0.0,
                 np.zeros((2, 1), dtype=np.int32))
    assert ei.value.code == 2 and "no CPU path" in str(ei.value)
    with pytest.raises(P.PipsortError):
        P.measure_fp64_peak()


def test_product_does_not_import_oracle():
    """Nothing under pipsort_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "pipsort_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(dp, fn)) as f:
                    txt = f.read()
                assert "oracle" not in txt.lower(), f"{fn} mentions oracle"


def test_argument_errors_before_cuda():
    import pipsort_b200 as P
    lib = P.lib()
    out = ctypes.c_void_p()
    assert lib.pipsort_create(None, 0, 0, ctypes.byref(out)) == 1
    assert b"null" in lib.pipsort_last_error()
    # three studies: the reference exits in log_prior (postcal.cpp:20-23)
    from pipsort_b200.engine import _Locus
    loc = _Locus(num_studies=3)
    assert lib.pipsort_create(ctypes.byref(loc), 0, 0, ctypes.byref(out)) == 3
    assert b"two studies" in lib.pipsort_last_error()


def test_new_entry_points_reject_bad_arguments_without_touching_cuda():
    import pipsort_b200 as P
    lib = P.lib()
    i32, u64 = ctypes.c_int32(), ctypes.c_uint64()
    assert lib.pipsort_sss(None, 3, 1000, ctypes.byref(i32), ctypes.byref(i32)) == 1
    assert lib.pipsort_sss_reset(None) == 1
    assert lib.pipsort_score_given_configs(None, None, 0, 0) == 1
    assert lib.pipsort_score_given_configs_device(None, None, 0, 0) == 1
    assert lib.pipsort_p2p_export(None, 2, None) == 1
    assert lib.pipsort_p2p_connect(None, None, 2, 0, 0) == 1
    assert lib.pipsort_p2p_reduce_to_root(None) == 1
    assert lib.pipsort_graph_begin(None) == 1
    assert lib.pipsort_graph_end(None, ctypes.byref(i32)) == 1
    assert lib.pipsort_graph_launch(None, 0) == 1
    assert lib.pipsort_prep_info_get(None, 0, None) == 1
    assert lib.pipsort_last_read_config_count(None, ctypes.byref(u64)) == 1
    assert lib.pipsort_posterior_exhaustive(None, 0, 0, 3, None, None) == 1
    assert lib.pipsort_posterior_exhaustive_batch(None, 2, 0, 0, 3, None, None) == 1
    assert lib.pipsort_posterior_exhaustive_batch(None, 0, 0, 0, 3, None, None) == 0
    assert lib.pipsort_preprocess_study(0, -1, None, None, None, None) == 1
    smap = np.zeros((2, 4), dtype=np.int32)
    b = (ctypes.c_uint64 * 3)()
    assert lib.pipsort_shard_ranks_for_map(smap.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 4, 2, 0, 0, b) == 1
    assert lib.pipsort_shard_ranks_for_map(smap.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 4, -1, 2, 0, b) == 1
    assert P.shard_ranks_for_map(smap, 2, 2) == [0, P.shard_ranks_for_map(smap, 2, 2)[1], 11]     # 1 + 4 + 6 ranks
    if not have_gpu():
        ld = np.eye(3)
        with pytest.raises(P.PipsortError) as ei:
            P.preprocess_study(ld, np.zeros(3))
        assert ei.value.code == 2 and "no CPU path" in str(ei.value)


def test_synth_counts():
    from pipsort_b200 import synth
    L = synth.make_locus(150)
    assert L.n_types() == (120, 30, 30) and L.U == 180
    assert synth.count_configs(L.snp_map, 3) == 12197751
    fl, nc = synth.flops_per_config_total(L.snp_map, 3)
    assert nc == 12197750 and abs(fl / nc - 68.8) < 0.1
    L2 = synth.make_locus(300)
    assert synth.count_configs(L2.snp_map, 2) == 352501           # config A
