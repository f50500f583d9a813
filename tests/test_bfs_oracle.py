"""Pins oracle/bfs_oracle.c with the reference's own BFS unit-test oracle (tests/unit/bfs_common.hpp: the
6-vertex graph and Floyd-Warshall distances with the bounds filter), plus a brute-force cross-check."""
import json
import os

import numpy as np

from _checkers import BfsOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bfs_unit_graph.json")
DIRS = {"OUT": 0, "IN": 1, "BOTH": 2}


def test_reference_unit_graph_all_directions_and_bounds():
    spec = json.load(open(GOLDEN))
    oracle = BfsOracle()
    e = np.array(spec["edges"], dtype=np.uint64)
    assert len(spec["cases"]) == 21
    for case in spec["cases"]:
        # the cursor's defaults: lower 1, upper unbounded (operator.cpp:2824-2828); the unit test's -1 means "not given"
        lower = 1 if case["lower"] == -1 else case["lower"]
        upper = 2**63 - 1 if case["upper"] == -1 else case["upper"]
        for source in range(spec["n"]):
            got = oracle.distances(spec["n"], e[:, 0], e[:, 1], source, DIRS[case["direction"]], lower, upper)
            assert got.tolist() == case["dist"][source], (case["direction"], case["lower"], case["upper"], source)


def test_bounds_that_produce_nothing():
    oracle = BfsOracle()
    assert (oracle.distances(3, [0, 1], [1, 2], 0, 0, 1, 0) == -1).all()   # upper < 1     (operator.cpp:2830)
    assert (oracle.distances(3, [0, 1], [1, 2], 0, 0, 3, 2) == -1).all()   # lower > upper
    assert oracle.distances(3, [0, 1], [1, 2], 0, 0, 0, 5).tolist() == [-1, 1, 2]  # lower 0 still hides the source


def test_random_graphs_against_networkx_style_bfs():
    oracle = BfsOracle()
    rng = np.random.default_rng(5)
    for n, m in [(50, 120), (300, 900)]:
        f, t = rng.integers(0, n, m), rng.integers(0, n, m)
        adj = [[] for _ in range(n)]
        for a, b in zip(f, t):
            adj[a].append(b)
        for source in [0, n // 2]:
            dist = [-1] * n
            dist[source] = 0
            frontier = [source]
            while frontier:
                nxt = []
                for u in frontier:
                    for v in adj[u]:
                        if dist[v] == -1:
                            dist[v] = dist[u] + 1
                            nxt.append(v)
                frontier = nxt
            dist[source] = -1
            assert oracle.distances(n, f, t, source).tolist() == dist
