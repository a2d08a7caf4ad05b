"""Parity of every launch shape of the exhaustive kernel with the CPU oracle (all accumulators, not only the total).

The planner (csrc/exhaustive.cuh, csrc/exh_plan.h) cuts the fixed sequence of warp-steps of a size class into chunks
whose cost it picks from the size of the rank range and of the GPU, so a small test locus only ever sees one shape.  These
tests force the other shapes (PIPSORT_EXH_CHUNK: chunks of 1, 2, 5 ... steps that end in the middle of a window, chunks that
span several x tiles, windows and first SNPs, one chunk for everything), run the BASELINE.json loci at full size against
the oracle on all host cores, and compare rank ranges of the saturating 1500-SNP locus with the oracle's walk over the
same ranks.
Reference loop being replaced: postcal.cpp:769-1044; accumulation :981-1030.
Tolerances: log-likelihoods 1e-10 relative, PIPs 1e-8 absolute (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import assert_results_match, engine_for, exh_plan_env, synth_as_oracle_locus, synth_locus

pytestmark = pytest.mark.gpu

_oracle_cache = {}


def oracle_range(key, L, c, lo=0, hi=None):
    from oracle import oracle as O
    k = (key, c, lo, hi)
    if k not in _oracle_cache:
        _oracle_cache[k] = O.exhaustive_omp_full(synth_as_oracle_locus(L), c, lo, hi)
    return _oracle_cache[k]


def test_b150c3_all_accumulators_whole_and_sharded():
    """BASELINE.json configs[3] (150 SNPs/study, U = 180, c = 3, 12,197,751 configurations): total, postValues, noCausal,
    sharedPips, sharedLL, notSharedLL against the oracle -- as one run, as the 2 / 4 / 8 work-weighted shards the multi-GPU
    path hands out (accumulated on one engine), and shard by shard (snp_map order, where the oracle can walk the same
    rank range)."""
    from pipsort_b200 import synth
    SL = synth_locus(150)
    want = oracle_range("B150", SL, 3)
    assert want.n_eval == synth.count_configs(SL.snp_map, 3) == 12197751
    with engine_for(SL, 3) as e:
        r = e.compute_total_likelihood(3)
        assert r.n_configs == want.n_eval
        assert_results_match(r, want)
        for parts in (2, 4, 8):
            b = e.shard_ranks(3, parts)
            e.reset()
            for i in range(parts):
                e.run_exhaustive(3, b[i], b[i + 1])
            rs = e.read()
            assert rs.n_configs == want.n_eval
            assert_results_match(rs, want)
    with engine_for(SL, 3, keep_order=True) as e:
        r = e.compute_total_likelihood(3)
        assert_results_match(r, want)
        b = e.shard_ranks(3, 8)
        for i in range(8):
            e.reset()
            e.run_exhaustive(3, b[i], b[i + 1])
            rs = e.read()
            ws = oracle_range("B150", SL, 3, b[i], b[i + 1])
            assert rs.n_configs == ws.n_eval
            assert_results_match(rs, ws)


CHUNKS = [1, 2, 5, 13, 33, 70, 300, 5000, 1 << 26]


@pytest.mark.parametrize("c", [2, 3])
@pytest.mark.parametrize("chunk", CHUNKS)
def test_forced_plan_shapes(chunk, c):
    """Every chunk shape, size classes 2 and 3, both SNP orders, whole and partial rank ranges.  70+70 SNPs at 50 %
    overlap: U = 105 = 4 x tiles (the lowest one partly empty: tiles are aligned to the top), all three SNP types, so
    chunks cut inside a segment, chunks over several tiles / windows / first SNPs, diagonal and partial tiles all occur."""
    from oracle import oracle as O
    SL = synth_locus(70, overlap=0.5, seed=77, sharing_param=0.25)
    U = SL.U
    assert U == 105
    tot = O.total_union_subsets(U, c)
    lowc = O.total_union_subsets(U, c - 1)
    whole = oracle_range("S70", SL, c)
    ranges = [(0, tot // 3), (lowc + 17, lowc + 17 + (tot - lowc) // 2), (tot - 4000, tot), (lowc - 5, lowc + 40)]
    with exh_plan_env(chunk):
        with engine_for(SL, c) as e:
            r = e.compute_total_likelihood(c)
            assert r.n_configs == whole.n_eval
            assert_results_match(r, whole)
        with engine_for(SL, c, keep_order=True) as e:
            r = e.compute_total_likelihood(c)
            assert r.n_configs == whole.n_eval
            assert_results_match(r, whole)
            for lo, hi in ranges:
                e.reset()
                e.run_exhaustive(c, lo, hi)
                rs = e.read()
                ws = oracle_range("S70", SL, c, lo, hi)
                assert rs.n_configs == ws.n_eval, (lo, hi)
                assert_results_match(rs, ws)


@pytest.mark.parametrize("overlap", [1.0, 0.0, 0.9, 0.3])
@pytest.mark.parametrize("chunk", [None, 7])
def test_type_layouts(overlap, chunk):
    """The step loops are specialised on what a (window, tile) segment cannot contain, and the x tiles end at the boundaries
    between SNP types (exh_plan.h, ExhTiles): loci whose type groups are missing (all SNPs shared / none shared), smaller
    than a tile (45 SNPs per study at 90 % overlap: groups of 40 + 5 + 5; at 30 %: 14 + 31 + 31) -- whole runs in both SNP
    orders and the sum over the work-weighted shards, size classes 2 and 3, against the oracle."""
    SL = synth_locus(45, overlap=overlap, seed=5 + int(10 * overlap), sharing_param=0.4)
    for c in (2, 3):
        whole = oracle_range(("T45", overlap), SL, c)
        with exh_plan_env(chunk):
            for keep in (False, True):
                with engine_for(SL, c, keep_order=keep) as e:
                    r = e.compute_total_likelihood(c)
                    assert r.n_configs == whole.n_eval
                    assert_results_match(r, whole)
                    b = e.shard_ranks(c, 5)
                    e.reset()
                    for i in range(5):
                        e.run_exhaustive(c, b[i], b[i + 1])
                    rs = e.read()
                    assert rs.n_configs == whole.n_eval
                    assert_results_match(rs, whole)


@pytest.mark.parametrize("forced", [None, 1400, 40])
def test_b1500c3_rank_ranges_against_oracle(forced):
    """The saturating locus of the roofline claim (1500 SNPs/study, U = 1800, c = 3; 1.2e10 configurations): rank ranges
    of ~1e6 union subsets at the start of size class 3, in its middle and at its end, in snp_map order, against the
    oracle's walk over the same ranks -- with the planner's own choice for the range, with the chunk size a whole-locus
    run uses (~1400 steps: dozens of x tiles per chunk) and with small chunks."""
    from oracle import oracle as O
    SL = synth_locus(1500)
    U = SL.U
    assert U == 1800
    tot = O.total_union_subsets(U, 3)
    low = O.total_union_subsets(U, 2)
    n = 1000000
    ranges = [(low - 1000, low + n), (low + (tot - low) // 2, low + (tot - low) // 2 + n), (tot - n, tot)]
    with exh_plan_env(forced):
        with engine_for(SL, 3, keep_order=True) as e:
            for lo, hi in ranges:
                e.reset()
                e.run_exhaustive(3, lo, hi)
                rs = e.read()
                ws = oracle_range("B1500", SL, 3, lo, hi)
                assert rs.n_configs == ws.n_eval, (lo, hi)
                assert_results_match(rs, ws)


def test_a300c2_forced_shapes():
    """BASELINE.json configs[2] (300 SNPs/study, c = 2) through small, medium and single-chunk plans as well."""
    SL = synth_locus(300, sharing_param=0.25)
    want = oracle_range("A300", SL, 2)
    for chunk in (3, 50, 1 << 26):
        with exh_plan_env(chunk):
            with engine_for(SL, 2) as e:
                r = e.compute_total_likelihood(2)
                assert r.n_configs == want.n_eval == 352501
                assert_results_match(r, want)
