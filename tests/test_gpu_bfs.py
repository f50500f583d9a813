"""-m gpu: the CUDA breadth-first expansion through the C ABI (include/mgb200_bfs.h) -- distances bit-exact
against the oracle and the reference's unit-test fixture (BASELINE config #5 at scale-24)."""
import json
import os

import numpy as np
import pytest

from _checkers import BfsOracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bfs_unit_graph.json")
DIRS = {"OUT": 0, "IN": 1, "BOTH": 2}


@pytest.fixture(scope="module")
def bfs():
    from memgraph_b200 import bfs
    return bfs


def test_reference_unit_graph(bfs):
    spec = json.load(open(GOLDEN))
    e = np.array(spec["edges"], dtype=np.uint64)
    with bfs.BfsGraph(spec["n"], e[:, 0], e[:, 1]) as g:
        for case in spec["cases"]:
            lower = 1 if case["lower"] == -1 else case["lower"]
            upper = bfs.INT64_MAX if case["upper"] == -1 else case["upper"]
            for source in range(spec["n"]):
                got, _ = g.distances(source, DIRS[case["direction"]], lower, upper)
                assert got.tolist() == case["dist"][source], (case, source)


@pytest.mark.parametrize("direction", [0, 1, 2])
def test_random_and_rmat_graphs_bit_exact(bfs, direction):
    import memgraph_b200 as mg
    oracle = BfsOracle()
    rng = np.random.default_rng(11)
    graphs = [(1, [], []), (5, [0, 0, 3], [0, 1, 4])]
    for n, m in [(1000, 3000), (20000, 200000)]:
        graphs.append((n, rng.integers(0, n, m), rng.integers(0, n, m)))
    f, t = mg.rmat_edges_host(16, 16 << 16)
    graphs.append((1 << 16, f, t))
    for n, f, t in graphs:
        with bfs.BfsGraph(n, f, t) as g:
            for source in {0, n // 3, n - 1}:
                for lower, upper in [(1, bfs.INT64_MAX), (2, 3), (1, 1)]:
                    got, st = g.distances(source, direction, lower, upper)
                    ref = oracle.distances(n, f, t, source, direction, lower, upper)
                    assert np.array_equal(got.astype(np.int64), ref), (n, source, direction, lower, upper)


def test_both_strategies_are_exercised_and_agree(bfs):
    """RMAT scale-20 from the biggest hub: the middle levels go bottom-up, the ends top-down."""
    import memgraph_b200 as mg
    oracle = BfsOracle()
    scale = 20
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m)
    with bfs.BfsGraph(n, f, t) as g:
        got, st = g.distances(0, 0)
        assert st["bottom_up_levels"] >= 1 and st["top_down_levels"] >= 1
        assert np.array_equal(got.astype(np.int64), oracle.distances(n, f, t, 0, 0))
        assert st["reached"] == int((got > 0).sum())
        assert st["edges_inspected"] < m  # direction optimisation skipped most of the edge list


def test_config5_rmat_scale24_distances(bfs):
    """BASELINE config #5: BFS expand on RMAT scale-24, distances bit-exact vs the (restated) reference."""
    import memgraph_b200 as mg
    oracle = BfsOracle()
    scale = int(os.environ.get("MGB200_BFS_TEST_SCALE", "24"))
    n, m = 1 << scale, 16 << scale
    f, t = mg.rmat_edges_host(scale, m)
    with bfs.BfsGraph(n, f, t) as g:
        for source in [0, 12345]:
            got, st = g.distances(source, 0)
            ref = oracle.distances(n, f, t, source, 0)
            assert np.array_equal(got.astype(np.int64), ref)
            print(f"scale-{scale} source {source}: levels {st['levels']} (td {st['top_down_levels']}, bu {st['bottom_up_levels']}), "
                  f"reached {st['reached']}, inspected {st['edges_inspected']} of {m} edges, {st['traverse_ms']:.3f} ms "
                  f"-> {m / st['traverse_ms'] / 1e6:.1f} GTEPS (input edges / traversal time)")
