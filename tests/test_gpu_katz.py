"""-m gpu: static Katz centrality on the device (include/mgb200_katz.h) against oracle/katz_oracle.cpp, which is pinned
bit-exact to the reference's katz.cpp (tests/test_katz_oracle.py).  Walk counts are integers, every other operation
keeps the reference's association and rounding: results and iteration counts must be IDENTICAL (np.array_equal) as
long as no walk count passes 2^53."""
import numpy as np
import pytest

from test_katz_oracle import oracle_katz

pytestmark = pytest.mark.gpu


def gpu_katz(n, f, t, alpha=0.2, eps=1e-2, guard=0):
    from memgraph_b200 import katz
    return katz.katz_from_edges(n, np.asarray(f, dtype=np.uint64), np.asarray(t, dtype=np.uint64), alpha, eps, guard)


def check(n, f, t, alpha=0.2, eps=1e-2):
    rc, ref, it = oracle_katz(n, f, t, alpha, eps)
    assert rc == 0
    got, st = gpu_katz(n, f, t, alpha, eps)
    assert st["iterations"] == it
    assert np.array_equal(got, ref)
    return st


def test_small_fixtures_bit_exact():
    check(4, [0, 1, 2], [1, 2, 3])                                # path
    check(5, [0, 1, 2, 3, 4], [1, 2, 3, 4, 0])                    # ring: all centralities tie
    check(6, [0, 0, 0, 0, 0], [1, 2, 3, 4, 5], alpha=0.1)         # out-star
    check(6, [1, 2, 3, 4, 5], [0, 0, 0, 0, 0], alpha=0.1)         # in-star
    check(4, [0, 0, 1, 1, 2, 2, 3, 3], [1, 1, 2, 2, 3, 3, 0, 0], alpha=0.1)  # multi-edges count twice
    check(3, [0, 1, 2], [0, 1, 2], alpha=0.3)                     # self-loops
    check(7, [0, 1], [1, 0])                                      # isolated vertices keep 0


def test_no_edges_and_empty_graph():
    got, st = gpu_katz(5, [], [])
    assert np.array_equal(got, np.zeros(5)) and st["iterations"] == 0
    got, st = gpu_katz(0, [], [])
    assert len(got) == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_graphs_bit_exact(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 3000))
    m = int(n * rng.integers(2, 6))
    f, t = rng.integers(0, n, m), rng.integers(0, n, m)
    deg_max = np.bincount(f, minlength=n).max()
    alpha = 0.5 / np.sqrt(float(deg_max) + 1.0)  # alpha^2 deg_max < 1: gamma positive, a real separation problem
    st = check(n, f, t, alpha=float(alpha), eps=1e-2)
    assert st["gamma"] > 0


def test_default_alpha_on_a_skewed_graph_stops_like_the_reference(monkeypatch):
    """alpha^2 deg_max > 1 makes gamma negative and the reference stop after ONE iteration; the drop-in must too.
    The heavy/SELL boundary is lowered so the heavy-row kernels take part."""
    import memgraph_b200 as mg
    monkeypatch.setenv("MGB200_HEAVY_MIN_DEGREE", "64")
    scale = 13
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m, seed=5)
    st = check(n, f, t)
    assert st["gamma"] < 0 and st["iterations"] == 1


def test_small_alpha_rmat_all_row_classes(monkeypatch):
    import memgraph_b200 as mg
    monkeypatch.setenv("MGB200_HEAVY_MIN_DEGREE", "64")
    scale = 12
    n, m = 1 << scale, 8 << scale
    f, t = mg.rmat_edges_host(scale, m, seed=9)
    deg_max = np.bincount(f.astype(np.int64), minlength=n).max()
    st = check(n, f, t, alpha=float(0.9 / np.sqrt(deg_max + 1.0)), eps=1e-3)
    assert st["iterations"] >= 2


def test_iteration_guard_reports_not_converged():
    from memgraph_b200 import katz
    f, t = [0, 1, 2, 3, 4], [1, 2, 3, 4, 0]
    rc, ref, it = oracle_katz(5, f, t, 0.2, 1e-9)
    if it > 1:
        with pytest.raises(katz.NotConverged):
            gpu_katz(5, f, t, 0.2, 1e-9, guard=1)


def test_pagerank_still_right_after_katz_on_the_same_handle():
    """Katz borrows the handle's contribution buffers and rank array; PageRank re-initialises all of it."""
    import memgraph_b200 as mg
    from memgraph_b200 import katz
    from _checkers import Oracle
    rng = np.random.default_rng(2)
    n, m = 2000, 12000
    f, t = rng.integers(0, n, m), rng.integers(0, n, m)
    ref, it = Oracle().pagerank(n, f, t, max_iterations=20, stop_epsilon=0.0)
    with mg.PageRankGraph.from_arrays(n, f, t) as g:
        before, _ = g.run(max_iterations=20, stop_epsilon=0.0)
        katz.set_katz(g, alpha=0.01, epsilon=1e-2)
        after, _ = g.run(max_iterations=20, stop_epsilon=0.0)
    assert np.array_equal(before, after)
    assert float(np.max(np.abs(after - ref) / ref)) < 1e-9
