"""katz_centrality.so (memgraph_b200/csrc/katz_centrality_module.cpp) on the in-memory mgp host.  No-GPU part: exports,
imports, registered signature, argument checks, loud failure without a device.  GPU part (-m gpu): rows against the
pinned oracle with the reference module's graph-view semantics (mg_utils.hpp:127-171, mg_graph.hpp:213-217)."""
import os
import subprocess

import numpy as np
import pytest

import _fakehost as fh
import conftest
from test_katz_oracle import oracle_katz

KATZ_MODULE_SO = os.path.join(fh.REPO, "memgraph_b200", "_build", "katz_centrality.so")
EXPECTED_SIGNATURE = ("get(alpha = 0.20000000000000001 :: FLOAT, epsilon = 0.01 :: FLOAT) :: "
                      "(node :: NODE, rank :: FLOAT)")


@pytest.fixture(scope="module")
def module():
    fh.host()
    with fh.Module(KATZ_MODULE_SO) as m:
        yield m


def test_exports_and_imports():
    fh.host()  # builds everything if needed
    out = subprocess.run(["nm", "-D", "--defined-only", KATZ_MODULE_SO], capture_output=True, text=True, check=True).stdout
    assert " T mgp_init_module" in out and " T mgp_shutdown_module" in out
    und = subprocess.run(["nm", "-D", "--undefined-only", KATZ_MODULE_SO], capture_output=True, text=True, check=True).stdout
    strong = {line.split()[-1] for line in und.splitlines() if " U mgp_" in line}
    # a subset of what the PageRank drop-in imports (no integer arguments here), nothing the host might lack
    assert strong <= set("""mgp_list_at mgp_value_get_double mgp_value_make_double mgp_value_make_vertex
        mgp_value_destroy mgp_type_float mgp_type_node mgp_module_add_read_procedure mgp_proc_add_opt_arg
        mgp_proc_add_result mgp_graph_approximate_vertex_count mgp_graph_approximate_edge_count mgp_graph_iter_vertices
        mgp_vertices_iterator_get mgp_vertices_iterator_next mgp_vertices_iterator_destroy mgp_vertex_iter_out_edges
        mgp_edges_iterator_get mgp_edges_iterator_next mgp_edges_iterator_destroy mgp_edge_get_to mgp_vertex_get_id
        mgp_graph_get_vertex_by_id mgp_graph_is_transactional mgp_result_new_record mgp_result_record_insert
        mgp_result_set_error_msg""".split())
    assert "cuda" not in und.lower() and "nccl" not in und.lower()


def test_signature_is_the_reference_signature(module):
    assert module.signature() == EXPECTED_SIGNATURE


def test_error_paths_without_device(module):
    before = fh.live_objects()
    with fh.Graph([5, 9, 2], [5, 9], [9, 2]) as g:
        with pytest.raises(fh.ProcedureError, match="must be of type FLOAT"):
            module.call(g, 1)
        if not conftest.HAVE_GPU:
            with pytest.raises(fh.ProcedureError, match=r"^katz_centrality\.get: CUDA error"):
                module.call(g)
    with fh.Graph([], [], []) as g:
        nodes, ranks = module.call(g)
        assert len(nodes) == 0 and len(ranks) == 0
    assert fh.live_objects() == before


def expected_rows(gids, src, dst, alpha, eps):
    order = np.sort(np.asarray(gids, dtype=np.int64))
    index = {int(g): i for i, g in enumerate(order)}
    f = np.array([index[int(s)] for s in src], dtype=np.uint64)
    t = np.array([index[int(d)] for d in dst], dtype=np.uint64)
    rc, ref, it = oracle_katz(len(order), f, t, alpha, eps)
    assert rc == 0
    return order, ref


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_rows_equal_the_oracle_bit_for_bit(module, seed):
    rng = np.random.default_rng(seed)
    n, m = 400, 1600
    gids = rng.choice(np.arange(10 * n, dtype=np.int64), size=n, replace=False)  # scattered, unsorted ids
    src, dst = gids[rng.integers(0, n, m)], gids[rng.integers(0, n, m)]
    deg_max = np.unique(src, return_counts=True)[1].max()
    alpha = float(0.5 / np.sqrt(deg_max + 1.0))
    order, ref = expected_rows(gids, src, dst, alpha, 1e-2)
    with fh.Graph(gids, src, dst) as g:
        nodes, ranks = module.call(g, alpha, 1e-2)
    assert np.array_equal(nodes, order) and np.array_equal(ranks, ref)


@pytest.mark.gpu
def test_defaults_and_analytical_mode(module):
    gids = np.array([3, 1, 2, 7], dtype=np.int64)
    src, dst = np.array([3, 1, 2], dtype=np.int64), np.array([1, 2, 7], dtype=np.int64)
    order, ref = expected_rows(gids, src, dst, 0.2, 1e-2)
    with fh.Graph(gids, src, dst) as g:
        nodes, ranks = module.call(g)
        assert np.array_equal(nodes, order) and np.array_equal(ranks, ref)
        g.hide_vertex(2)
        with pytest.raises(fh.ProcedureError, match=r"^katz_centrality\.get: Invalid ID!$"):
            module.call(g)
    with fh.Graph(gids, src, dst, transactional=False) as g:
        g.hide_vertex(2)
        nodes, ranks = module.call(g)
        keep = order != 2
        assert np.array_equal(nodes, order[keep]) and np.array_equal(ranks, ref[keep])
