"""gpu_bfs query module through the fake mgp host.  No-GPU part: load, signature, argument validation, loud
failure without a device.  GPU part (-m gpu): distances equal the oracle's and the reference unit fixture."""
import json
import os

import numpy as np
import pytest

import _fakehost as fh
import conftest
from _checkers import BfsOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bfs_unit_graph.json")
SIGNATURE = ("distances(source :: NODE, direction = 0 :: INTEGER, lower_bound = 1 :: INTEGER, "
             "upper_bound = -1 :: INTEGER) :: (node :: NODE, distance :: INTEGER)")


@pytest.fixture(scope="module")
def module():
    m = fh.Module(fh.BFS_MODULE_SO)
    yield m
    assert m.close() == 0


def test_signature_and_argument_validation(module):
    assert module.signature("distances") == SIGNATURE
    with fh.Graph([1, 2, 3], [1, 2], [2, 3]) as g:
        with pytest.raises(fh.ProcedureError, match="requires between 1 and 4 arguments"):
            module.call(g, proc="distances")
        with pytest.raises(fh.ProcedureError, match="must be of type NODE"):
            module.call(g, 1, proc="distances")
        with pytest.raises(fh.ProcedureError, match="must be of type INTEGER"):
            module.call(g, fh.Node(1), 0.5, proc="distances")
        with pytest.raises(fh.ProcedureError, match=r"^gpu_bfs\.distances: direction must be 0 \(OUT\), 1 \(IN\) or 2 \(BOTH\)$"):
            module.call(g, fh.Node(1), 7, proc="distances")
        if not conftest.HAVE_GPU:
            before = fh.live_objects()
            with pytest.raises(fh.ProcedureError, match=r"^gpu_bfs\.distances: CUDA error"):
                module.call(g, fh.Node(1), proc="distances")
            assert fh.live_objects() == before


@pytest.mark.gpu
def test_reference_unit_fixture_through_the_module(module):
    spec = json.load(open(GOLDEN))
    gids = [100 + 7 * v for v in range(spec["n"])]
    src = [gids[a] for a, _ in spec["edges"]]
    dst = [gids[b] for _, b in spec["edges"]]
    dirs = {"OUT": 0, "IN": 1, "BOTH": 2}
    with fh.Graph(gids, src, dst) as g:
        for case in spec["cases"]:
            lower = 1 if case["lower"] == -1 else case["lower"]
            for source in range(spec["n"]):
                nodes, dist = module.call(g, fh.Node(gids[source]), dirs[case["direction"]], lower, case["upper"],
                                          proc="distances")
                got = {int(n): int(d) for n, d in zip(nodes, dist)}
                exp = {gids[v]: d for v, d in enumerate(case["dist"][source]) if d != -1}
                assert got == exp, (case["direction"], case["lower"], case["upper"], source)


@pytest.mark.gpu
def test_random_graph_through_the_module(module):
    oracle = BfsOracle()
    rng = np.random.default_rng(21)
    n, m = 5000, 30000
    gids = rng.choice(np.arange(10 * n, dtype=np.int64), size=n, replace=False)
    order = np.sort(gids)
    s, t = rng.integers(0, n, m), rng.integers(0, n, m)
    before = fh.live_objects()
    with fh.Graph(gids, gids[s], gids[t]) as g:
        for source_dense in [0, n // 2]:
            for direction in [0, 1, 2]:
                nodes, dist = module.call(g, fh.Node(order[source_dense]), direction, proc="distances")
                # dense ids = ascending-gid visit order (PullGraph), like the PageRank module
                dense_of = {int(x): i for i, x in enumerate(order)}
                f = np.array([dense_of[int(x)] for x in gids[s]])
                to = np.array([dense_of[int(x)] for x in gids[t]])
                ref = oracle.distances(n, f, to, source_dense, direction)
                got = np.full(n, -1, dtype=np.int64)
                got[[dense_of[int(x)] for x in nodes]] = dist.astype(np.int64)
                assert np.array_equal(got, ref), (source_dense, direction)
    assert fh.live_objects() == before
