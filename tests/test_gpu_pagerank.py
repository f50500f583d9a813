"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle
(oracle/pagerank_oracle.c, pinned bit-exact to the reference) and the committed golden fixtures.

Tolerance: north_star demands every rank within 1e-6 RELATIVE of the reference after the same
iteration count.  FP64 end to end leaves ~6 orders of margin, so the tests assert 1e-9 relative
(only the order of additions inside a row differs from the reference)."""
import ctypes
import json
import os

import numpy as np
import pytest

from _checkers import Oracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-9          # asserted
NORTH_STAR_REL = 1e-6   # the contract


@pytest.fixture(scope="module")
def mg():
    import memgraph_b200
    return memgraph_b200


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def rel_err(got, ref):
    ref = np.asarray(ref)
    return float(np.max(np.abs(np.asarray(got) - ref) / ref)) if len(ref) else 0.0


def gpu_pagerank(mg, n, frm, to, **kw):
    with mg.PageRankGraph.from_arrays(n, frm, to) as g:
        ranks, st = g.run(**kw)
    return ranks, st


def test_unit_golden_vectors(mg, oracle):
    """pagerank_test.cpp:35-65 through the C ABI: reference tolerance (1e-3) and tight vs the oracle."""
    spec = json.load(open(os.path.join(GOLDEN, "pagerank_unit_vectors.json")))
    for case in spec["cases"]:
        edges = np.array(case["edges"], dtype=np.uint64).reshape(-1, 2)
        g = mg.PageRankGraph(case["n"], case["m"], edges)
        ranks = mg.parallel_iterative_pagerank(g)  # default arguments, like the reference test
        g.close()
        exp = np.array(case["expected"])
        assert len(ranks) == len(exp)
        if len(exp):
            err = np.abs(ranks - exp)
            assert err.max() < 1e-3 and err.mean() < 1e-3, case  # mg_test_utils.hpp:107-112
            ref, _ = oracle.pagerank(case["n"], edges[:, 0], edges[:, 1], **spec["args"])
            assert rel_err(ranks, ref) < REL_TOL, case


def test_committed_reference_outputs(mg):
    """Outputs of the real reference algorithm (tests/golden/make_golden.py) -- no oracle involved."""
    spec = json.load(open(os.path.join(GOLDEN, "pagerank_ref_outputs.json")))
    for case in spec["cases"]:
        a = dict(case["args"])
        threads = a.pop("num_of_threads")
        ranks, it = mg.pagerank_from_edges(case["n"], case["from"], case["to"], number_of_threads=threads, **a)
        exp = np.array([float(x) for x in case["ranks"]])
        assert rel_err(ranks, exp) < REL_TOL, (case["n"], case["args"])


def test_edge_case_semantics(mg, oracle):
    """SURVEY 8a edge-case table, each row."""
    # num_of_threads == 0 -> the reference's error text
    with pytest.raises(mg.MgB200Error, match=r"Number of threads can't be zero \(0\)!"):
        mg.pagerank_from_edges(2, [0], [1], number_of_threads=0)
    g = mg.PageRankGraph(2, 1, [[0, 1]])
    with pytest.raises(mg.MgB200Error, match=r"Number of threads can't be zero \(0\)!"):
        mg.parallel_iterative_pagerank(g, number_of_threads=0)
    # negative thread count: wraps, result unaffected
    a = mg.parallel_iterative_pagerank(g, number_of_threads=-1)
    b = mg.parallel_iterative_pagerank(g, number_of_threads=1)
    assert np.array_equal(a, b)
    g.close()
    # empty graph: empty vector, one nominal iteration
    r, it = mg.pagerank_from_edges(0, [], [])
    assert len(r) == 0 and it == 1
    # max_iterations == 0: normalised uniform vector
    r, it = mg.pagerank_from_edges(4, [0, 1], [1, 2], max_iterations=0)
    assert it == 0 and np.array_equal(r, np.full(4, 0.25))
    # max_iterations < 0 wraps to "until converged"
    r, it = mg.pagerank_from_edges(5, [0, 1, 2, 3], [1, 2, 3, 4], max_iterations=-1, stop_epsilon=1e-12)
    ref, rit = oracle.pagerank(5, [0, 1, 2, 3], [1, 2, 3, 4], max_iterations=-1, stop_epsilon=1e-12)
    assert it == rit and rel_err(r, ref) < REL_TOL
    # damping 1.0 with dangling nodes: all mass leaks -> NaN
    r, _ = mg.pagerank_from_edges(3, [0, 1], [1, 2], damping_factor=1.0, max_iterations=50)
    assert np.all(np.isnan(r))
    # isolated vertices are kept
    r, _ = mg.pagerank_from_edges(3, [0], [1], max_iterations=20, stop_epsilon=0.0)
    ref, _ = oracle.pagerank(3, [0], [1], max_iterations=20, stop_epsilon=0.0)
    assert rel_err(r, ref) < REL_TOL and r[2] > 0
    # out-of-range endpoint is rejected (the reference would read out of bounds)
    with pytest.raises(mg.MgB200Error, match="out of range"):
        mg.pagerank_from_edges(3, [0, 5], [1, 2])
    # single vertex, self loops, multi-edges
    for n, f, t in [(1, [], []), (1, [0], [0]), (2, [0, 0], [1, 1]), (2, [1], [1])]:
        r, it = mg.pagerank_from_edges(n, f, t)
        ref, rit = oracle.pagerank(n, f, t)
        assert it == rit and rel_err(r, ref) < REL_TOL


def erdos_renyi(n, m, seed=42):
    rng = np.random.default_rng(seed)  # uniform pairs with replacement, self-loops / multi-edges allowed
    return rng.integers(0, n, size=m, dtype=np.uint64), rng.integers(0, n, size=m, dtype=np.uint64)


@pytest.mark.parametrize("kw", [dict(max_iterations=100, damping_factor=0.85, stop_epsilon=1e-5),
                                dict(max_iterations=20, damping_factor=0.85, stop_epsilon=0.0),
                                dict(max_iterations=1000, damping_factor=0.5, stop_epsilon=1e-12)])
def test_config1_erdos_renyi_10k_50k(mg, oracle, kw):
    """BASELINE config #1 graph on the GPU: ranks and executed-iteration count equal the oracle's."""
    n, m = 10_000, 50_000
    f, t = erdos_renyi(n, m)
    ranks, st = gpu_pagerank(mg, n, f, t, **kw)
    ref, it = oracle.pagerank(n, f, t, **kw)
    assert st.iterations == it
    assert rel_err(ranks, ref) < REL_TOL
    assert abs(ranks.sum() - 1.0) < 1e-12


@pytest.mark.parametrize("heavy_min,seg", [(None, None), (8, 32), (2, 3), (64, 4096)])
def test_rmat_small_all_row_classes(mg, oracle, monkeypatch, heavy_min, seg):
    """RMAT scale-14 with the heavy/SELL boundary moved around so every kernel (segments, finish,
    SELL, zero rows) carries a real share of the rows."""
    if heavy_min is not None:
        monkeypatch.setenv("MGB200_HEAVY_MIN_DEGREE", str(heavy_min))
        monkeypatch.setenv("MGB200_SEGMENT_EDGES", str(seg))
    scale = 14
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m)
    with mg.PageRankGraph.from_arrays(n, f, t) as g:
        info = dict(g.info)
        ranks, st = g.run(max_iterations=20, stop_epsilon=0.0)
        again, _ = g.run(max_iterations=20, stop_epsilon=0.0)
    assert info["heavy_rows"] + info["sell_rows"] + info["zero_rows"] == n
    assert info["local_edges"] == m
    if heavy_min is not None and heavy_min <= 8:
        assert info["heavy_rows"] > 0 and info["heavy_segments"] >= info["heavy_rows"]
    ref, it = oracle.pagerank(n, f, t, max_iterations=20, stop_epsilon=0.0)
    assert st.iterations == it == 20
    assert rel_err(ranks, ref) < REL_TOL
    assert np.array_equal(ranks, again)  # deterministic: fixed summation order, no float atomics


@pytest.mark.parametrize("env", [{"MGB200_IDX_FLAGS": "1"},
                                 {"MGB200_IDX_FLAGS": "0", "MGB200_FORCE_MULTI_PATH": "1"},
                                 {"MGB200_L1_HOT_K": "-1"}, {"MGB200_L2_HOT_MB": "0", "MGB200_IDX_FLAGS": "1"},
                                 {"MGB200_SELL_KERNEL": "stream"},  # TMA index ring + LDGSTS gathers (sell_stream.cuh)
                                 {"MGB200_SELL_MODE": "0", "MGB200_SELL_WORK_ITEMS": "37"},   # ticket queue
                                 {"MGB200_SELL_MODE": "1", "MGB200_SELL_WORK_ITEMS": "5"},    # static deal
                                 {"MGB200_SELL_MODE": "0", "MGB200_SELL_WORK_ITEMS": "100000"},
                                 {"MGB200_OVERLAP_EPILOGUE": "0"},
                                 {"MGB200_SMEM_TABLE_KB": "0"}, {"MGB200_SMEM_TABLE_KB": "1"}, {"MGB200_SMEM_TABLE_KB": "64"},
                                 {"MGB200_SMEM_TABLE_KB": "200"},                              # hot table sizes (TMA fill)
                                 {"MGB200_SMEM_TABLE_KB": "96", "MGB200_IDX_FLAGS": "1"}])
def test_gather_path_variants_are_bit_identical(mg, oracle, monkeypatch, env):
    """How a gather learns its cache policy (range policy / hot flags baked into the stored index / per-gather owner
    lookup, with or without L1 hints), which kernel walks the SELL slices (LDG rows kernel / TMA + LDGSTS stream kernel)
    and how the slices are handed to warps (ticket queue / static deal, any item count) must never change a single bit
    of the result: the flags live in index bits 31/30 and are masked off before addressing, and every row is summed by
    one lane in ascending source order whatever warp gets it."""
    scale = 13
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m, seed=7)
    monkeypatch.setenv("MGB200_HEAVY_MIN_DEGREE", "64")  # a real heavy class at this size
    base, st0 = gpu_pagerank(mg, n, f, t, max_iterations=20, stop_epsilon=0.0)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got, st1 = gpu_pagerank(mg, n, f, t, max_iterations=20, stop_epsilon=0.0)
    assert st0.iterations == st1.iterations == 20
    assert np.array_equal(base, got)
    ref, _ = oracle.pagerank(n, f, t, max_iterations=20, stop_epsilon=0.0)
    assert rel_err(got, ref) < REL_TOL


def test_rmat_generator_host_equals_device(mg):
    from memgraph_b200 import _native as N
    lib = N.lib()
    scale, m = 18, 1 << 20
    d_f, d_t = N.vp(), N.vp()
    assert lib.mgb200_device_malloc(0, 4 * m, ctypes.byref(d_f)) == 0
    assert lib.mgb200_device_malloc(0, 4 * m, ctypes.byref(d_t)) == 0
    mg.rmat_edges_device(scale, m, d_f, d_t, seed=42, first_edge=12345)
    hf = np.empty(m, dtype=np.uint32)
    ht = np.empty(m, dtype=np.uint32)
    assert lib.mgb200_copy_to_host(0, hf.ctypes.data, d_f, 4 * m) == 0
    assert lib.mgb200_copy_to_host(0, ht.ctypes.data, d_t, 4 * m) == 0
    lib.mgb200_device_free(0, d_f)
    lib.mgb200_device_free(0, d_t)
    f, t = mg.rmat_edges_host(scale, m, seed=42, first_edge=12345)
    assert np.array_equal(hf, f.astype(np.uint32)) and np.array_equal(ht, t.astype(np.uint32))
    # quadrant marginals: P(src bit = 1) = c + d = 0.24, P(dst bit = 1) = b + d = 0.24
    assert abs(float(((f >> (scale - 1)) & 1).mean()) - 0.24) < 0.005
    assert abs(float(((t >> (scale - 1)) & 1).mean()) - 0.24) < 0.005


def test_convergence_iteration_count_default_eps(mg, oracle):
    """Default stop_epsilon: the L-infinity absolute test on un-normalised ranks must stop at exactly
    the reference's iteration (pagerank.cpp:138-150)."""
    scale = 16
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m)
    ranks, st = gpu_pagerank(mg, n, f, t)  # (100, 0.85, 1e-5)
    ref, it = oracle.pagerank(n, f, t)
    assert st.iterations == it and 1 < it < 100
    assert rel_err(ranks, ref) < REL_TOL


def test_config2_rmat_scale22(mg, oracle):
    """BASELINE config #2: RMAT scale-22, 20 iterations, ranks vs the oracle within 1e-6 relative
    (asserted at 1e-9)."""
    scale = int(os.environ.get("MGB200_TEST_SCALE", "22"))
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m)
    ranks, st = gpu_pagerank(mg, n, f, t, max_iterations=20, stop_epsilon=0.0)
    ref, it = oracle.pagerank(n, f, t, max_iterations=20, stop_epsilon=0.0)
    assert st.iterations == it == 20
    err = rel_err(ranks, ref)
    print(f"scale-{scale}: max relative error vs oracle {err:.3e}, max abs {np.max(np.abs(ranks - ref)):.3e}, "
          f"sum {ranks.sum():.15f}")
    assert err < REL_TOL < NORTH_STAR_REL
    assert abs(ranks.sum() - 1.0) < 1e-12


def test_full_size_properties_scale26(mg):
    """BASELINE config #3 size (RMAT scale-26, ~1.07 B edges): size-independent properties --
    sum to 1, positivity with the (1-d)/N floor, bit-reproducibility, and agreement of one extra
    iteration with a device-independent recomputation on a vertex sample is covered at scale-22."""
    from memgraph_b200 import _native as N
    lib = N.lib()
    total = ctypes.c_size_t(0)
    assert lib.mgb200_device_info(0, None, 0, None, ctypes.byref(total)) == 0
    scale = int(os.environ.get("MGB200_FULL_SCALE", "26"))
    if total.value < 60 * 2**30:
        scale = min(scale, 22)
    n, m = 1 << scale, 16 << scale
    d_f, d_t = N.vp(), N.vp()
    assert lib.mgb200_device_malloc(0, 4 * m, ctypes.byref(d_f)) == 0
    assert lib.mgb200_device_malloc(0, 4 * m, ctypes.byref(d_t)) == 0
    mg.rmat_edges_device(scale, m, d_f, d_t)
    g = mg.PageRankGraph.from_device(n, m, d_f, d_t)
    lib.mgb200_device_free(0, d_f)
    lib.mgb200_device_free(0, d_t)
    info = g.info
    assert info["heavy_rows"] + info["sell_rows"] + info["zero_rows"] == n and info["local_edges"] == m
    r1, st1 = g.run(max_iterations=20, stop_epsilon=0.0)
    r2, st2 = g.run(max_iterations=20, stop_epsilon=0.0)
    g.close()
    assert st1.iterations == st2.iterations == 20
    assert np.array_equal(r1, r2)
    assert abs(r1.sum() - 1.0) < 1e-11
    floor = (1 - 0.85) / n / st1.rank_sum
    assert r1.min() >= floor * (1 - 1e-12)
    # vertices with in-degree 0 sit exactly on the floor; RMAT leaves a large fraction there
    assert np.count_nonzero(r1 == r1.min()) == info["zero_rows"]


def test_config3_rmat_scale26_ranks_vs_oracle(mg):
    """BASELINE config #3 -- the headline size: all 2^26 ranks of the CUDA path against the CPU checker on the same
    RMAT scale-26 bytes (oracle/rmat_oracle.c == csrc/rmat.hpp, pinned byte for byte in test_oracle.py), 20 iterations,
    stop_epsilon = 0.  The checker is oracle/pagerank_pull_oracle.c (multi-threaded pull restatement, itself pinned to
    1e-12 on the bit-exact reference restatement -- the ladder SURVEY 8c prescribes for this size; the sequential
    restatement needs ~8 minutes here).  Two numbers: the plain relative error (north_star bar 1e-6; asserted 1e-8), which
    at this size is dominated by ONE scalar -- the reference normalises by a sequential std::accumulate over 67 M addends,
    the device by a tree sum -- and the error after dividing that scalar out (asserted 1e-11): the per-vertex arithmetic."""
    from memgraph_b200 import _native as N
    from _checkers import oracle_rmat_edges, pull_oracle_pagerank
    lib = N.lib()
    total = ctypes.c_size_t(0)
    assert lib.mgb200_device_info(0, None, 0, None, ctypes.byref(total)) == 0
    scale = int(os.environ.get("MGB200_FULL_SCALE", "26"))
    host_gb = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2**30
    if total.value < 60 * 2**30 or host_gb < 96:
        scale = min(scale, 23)  # small box: same test, smaller graph
    n, m = 1 << scale, 16 << scale
    d_f, d_t = N.vp(), N.vp()
    assert lib.mgb200_device_malloc(0, 4 * m, ctypes.byref(d_f)) == 0
    assert lib.mgb200_device_malloc(0, 4 * m, ctypes.byref(d_t)) == 0
    mg.rmat_edges_device(scale, m, d_f, d_t)
    with mg.PageRankGraph.from_device(n, m, d_f, d_t) as g:
        lib.mgb200_device_free(0, d_f)
        lib.mgb200_device_free(0, d_t)
        ranks, st = g.run(max_iterations=20, stop_epsilon=0.0)
    f, t = oracle_rmat_edges(scale, m)
    ref, it = pull_oracle_pagerank(n, f, t, max_iterations=20, damping_factor=0.85, stop_epsilon=0.0)
    del f, t
    assert st.iterations == it == 20
    ratio = ranks / ref
    err = float(np.max(np.abs(ratio - 1.0)))
    scalar = float(np.median(ratio))
    err_shape = float(np.max(np.abs(ratio / scalar - 1.0)))
    print(f"scale-{scale}: max relative error vs oracle {err:.3e} (normalisation scalar {scalar - 1.0:+.3e}, "
          f"per-vertex {err_shape:.3e}), sum {ranks.sum():.15f}")
    assert err < 1e-8 < NORTH_STAR_REL
    assert err_shape < 1e-11


def test_many_small_random_graphs(mg, oracle):
    """200 seeded random multigraphs (N 1..60, self-loops, parallel edges, isolated vertices, sink-only graphs):
    ranks and executed-iteration counts equal the oracle's for the default arguments and for a fixed count."""
    rng = np.random.default_rng(20260921)
    mismatched = []
    for trial in range(200):
        n = int(rng.integers(1, 61))
        m = int(rng.integers(0, 4 * n + 1))
        f, t = rng.integers(0, n, m), rng.integers(0, n, m)
        if trial % 7 == 0 and m:
            t[:] = t[0]  # a star: one hub destination, many dangling sources
        kw = dict() if trial % 2 else dict(max_iterations=int(rng.integers(1, 30)), damping_factor=float(rng.uniform(0.05, 0.99)),
                                           stop_epsilon=0.0)
        ranks, it = mg.pagerank_from_edges(n, f, t, **kw)
        ref, rit = oracle.pagerank(n, f, t, **kw)
        if it != rit:
            # Only legitimate cause: stop_epsilon == 0 ends the loop when two successive vectors are BIT-identical,
            # which depends on the order of additions inside a row (the reference's own count changes with its
            # thread count).  Tiny graphs sit on an exact fixed point up to rounding noise: one side sees two
            # identical vectors an iteration or two before the other (which may dither in the last bit until the
            # cap).  The ranks still agree to rounding (asserted below).  With a positive epsilon the counts must agree.
            assert kw.get("stop_epsilon", 1e-5) == 0.0 and min(it, rit) < kw["max_iterations"], \
                (trial, n, m, kw, it, rit)
            # ... and it must be THAT cause: once one side has stopped, the other side's vector is the same fixed point up
            # to the last bits (a real divergence would move the ranks), and the reference itself must be order-sensitive
            # on this graph or have stopped within the dithering window
            counts = {oracle.pagerank(n, f, t, **dict(kw, num_of_threads=T))[1] for T in (1, 2, 3, 8)}
            few = oracle.pagerank(n, f, t, **dict(kw, max_iterations=min(it, rit)))[0]
            assert len(counts) > 1 or rel_err(few, ref) < 1e-14, (trial, n, m, kw, it, rit, counts)
            mismatched.append((trial, it, rit, sorted(counts)))
        assert rel_err(ranks, ref) < REL_TOL, (trial, n, m, kw)
    print(f"iteration-count mismatches under stop_epsilon=0: {mismatched}")
    assert len(mismatched) <= 10  # a handful of exact-fixed-point graphs, not a pattern


def test_abort_hook_single_gpu_and_handle_reuse(mg, oracle):
    """should_abort (mgp_must_abort) is polled between 32-iteration batches; the request travels through the device
    (iter_end_kernel) and comes back as MGB200_ERR_ABORTED; the handle then runs a normal call correctly."""
    from memgraph_b200 import _native as N
    n, m = 5000, 40000
    rng = np.random.default_rng(5)
    f, t = rng.integers(0, n, m), rng.integers(0, n, m)
    polls = []
    with mg.PageRankGraph.from_arrays(n, f, t) as g:
        with pytest.raises(mg.MgB200Error) as ei:
            g.run(max_iterations=10**6, stop_epsilon=-1.0, should_abort=lambda: polls.append(1) or len(polls) >= 2)
        assert ei.value.code == N.ERR_ABORTED and len(polls) == 2
        ranks, st = g.run(max_iterations=20, stop_epsilon=0.0)
    ref, it = oracle.pagerank(n, f, t, max_iterations=20, stop_epsilon=0.0)
    assert st.iterations == it and rel_err(ranks, ref) < REL_TOL


def test_all_nan_deltas_stop_like_the_reference(mg, oracle):
    """CheckContinueIterate (:138-150) continues iff SOME element has |delta| > eps.  With a NaN damping factor every
    delta is NaN: no element qualifies, the reference stops after one iteration even for a negative epsilon -- a running
    maximum that starts at 0 would see 0 > eps and go on."""
    n, m = 300, 2000
    rng = np.random.default_rng(11)
    f, t = rng.integers(0, n, m), rng.integers(0, n, m)
    for eps in (-1.0, 0.0, 1e-5):
        ranks, it = mg.pagerank_from_edges(n, f, t, max_iterations=50, damping_factor=float("nan"), stop_epsilon=eps)
        ref, rit = oracle.pagerank(n, f, t, max_iterations=50, damping_factor=float("nan"), stop_epsilon=eps)
        assert it == rit == 1, (eps, it, rit)
        assert np.isnan(ranks).all() and np.isnan(ref).all()
    # negative epsilon with ordinary numbers: every delta >= 0 > eps, the loop runs to the cap on both sides
    ranks, it = mg.pagerank_from_edges(n, f, t, max_iterations=37, stop_epsilon=-1.0)
    ref, rit = oracle.pagerank(n, f, t, max_iterations=37, stop_epsilon=-1.0)
    assert it == rit == 37 and rel_err(ranks, ref) < REL_TOL


def test_streamed_and_chunked_builds_are_bit_identical(mg, monkeypatch):
    """The build walks its edge source a chunk at a time (core.hpp EdgeSource): resident arrays in one chunk, in many
    small chunks, and the RMAT generator stream (no COO materialised) must give the same layout and the same bits."""
    scale = 14
    n, m = 1 << scale, 16 << scale
    monkeypatch.setenv("MGB200_HEAVY_MIN_DEGREE", "64")
    f, t = mg.rmat_edges_host(scale, m, seed=11)
    base, st = gpu_pagerank(mg, n, f, t, max_iterations=20, stop_epsilon=0.0)
    monkeypatch.setenv("MGB200_BUILD_CHUNK_EDGES", "7777")
    chunked, _ = gpu_pagerank(mg, n, f, t, max_iterations=20, stop_epsilon=0.0)
    with mg.PageRankGraph.from_rmat(scale, m, seed=11) as g:
        streamed, _ = g.run(max_iterations=20, stop_epsilon=0.0)
        info = dict(g.info)
    assert np.array_equal(base, chunked) and np.array_equal(base, streamed)
    assert info["local_edges"] == m and info["heavy_rows"] > 0
