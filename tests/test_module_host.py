"""No-GPU tests of the drop-in boundary: the fake mgp host is validated by loading the REFERENCE's own
pagerank module (BASELINE config #1), then the B200 module is checked for everything that does not
need a device: exports, registered signature, argument validation, error texts, loud failure
without a GPU, no leaked host objects."""
import json
import os
import subprocess

import numpy as np
import pytest

import _fakehost as fh
from _checkers import Oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF_MODULE = os.path.exists(fh.REF_MODULE_SO)
needs_ref_module = pytest.mark.skipif(not HAVE_REF_MODULE, reason="oracle/_ref/pagerank_reference.so not built")

EXPECTED_SIGNATURE = ("get(max_iterations = 100 :: INTEGER, damping_factor = 0.84999999999999998 :: FLOAT, "
                      "stop_epsilon = 1.0000000000000001e-05 :: FLOAT, num_of_threads = 1 :: INTEGER) :: "
                      "(node :: NODE, rank :: FLOAT)")


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def ref_module():
    m = fh.Module(fh.REF_MODULE_SO)
    yield m
    m.close()


@pytest.fixture(scope="module")
def b200_module():
    m = fh.Module(fh.MODULE_SO)
    yield m
    assert m.close() == 0


def scattered_graph(n, m, seed):
    """Non-contiguous, unsorted gids, like a real database hands out."""
    rng = np.random.default_rng(seed)
    gids = rng.choice(np.arange(10 * n + 7, dtype=np.int64), size=n, replace=False)
    s = rng.integers(0, n, size=m)
    t = rng.integers(0, n, size=m)
    return gids, gids[s], gids[t]


def oracle_through_module_semantics(oracle, gids, src, dst, **kw):
    """What pagerank.get must return: dense ids in ascending-gid visit order, one row per vertex."""
    order = np.sort(gids)
    frm, to = oracle.map_gids(order, src, dst)
    # the module sees each vertex's out-edges grouped by source in visit order
    perm = np.argsort(frm, kind="stable")
    ranks, it = oracle.pagerank(len(order), frm[perm], to[perm], **kw)
    return order, ranks, it


# ---- the host itself, validated with the reference's own module (config #1) ---------------------------------

@needs_ref_module
def test_reference_module_signature(ref_module):
    assert ref_module.signature() == EXPECTED_SIGNATURE


@needs_ref_module
def test_config1_reference_module_erdos_renyi_10k_50k(ref_module, oracle):
    """BASELINE config #1: 10k-node / 50k-edge Erdos-Renyi graph through the reference algorithm and
    the C-ABI module loader, no GPU.  Default args and (20, 0.85, 0.0, 8)."""
    n, m = 10_000, 50_000
    gids, src, dst = scattered_graph(n, m, seed=42)
    before = fh.live_objects()
    with fh.Graph(gids, src, dst) as g:
        nodes, ranks = ref_module.call(g)
        nodes2, ranks2 = ref_module.call(g, 20, 0.85, 0.0, 8)
    assert fh.live_objects() == before  # the reference module leaks nothing through this host
    order, exp, _ = oracle_through_module_semantics(oracle, gids, src, dst)
    assert np.array_equal(nodes, order)
    assert np.array_equal(ranks, exp)  # same algorithm, same order of operations, T = 1
    _, exp2, _ = oracle_through_module_semantics(oracle, gids, src, dst, max_iterations=20, stop_epsilon=0.0,
                                                 num_of_threads=8)
    assert np.array_equal(nodes2, order) and np.array_equal(ranks2, exp2)


def run_e2e_fixture(module, case):
    ids = sorted({v for e in case["edges_by_node_id"] for v in e})
    gid_of = {node_id: 1000 + 3 * i for i, node_id in enumerate(ids)}  # gids follow creation order here
    src = [gid_of[a] for a, _ in case["edges_by_node_id"]]
    dst = [gid_of[b] for _, b in case["edges_by_node_id"]]
    with fh.Graph(list(gid_of.values()), src, dst) as g:
        nodes, ranks = module.call(g, *case["call_args"])
    id_of = {v: k for k, v in gid_of.items()}
    rows = sorted((id_of[int(gid)], float(r)) for gid, r in zip(nodes, ranks))  # ORDER BY node ASC
    return rows


@needs_ref_module
def test_reference_module_e2e_fixtures(ref_module):
    """mage/tests/e2e/pagerank_test/*: expected rows at abs 1e-3 (test_module.py:21,76)."""
    spec = json.load(open(os.path.join(GOLDEN, "pagerank_e2e_cases.json")))
    assert {c["name"] for c in spec["cases"]} == {"test_chain", "test_empty", "test_influential_node",
                                                  "test_influential_node_transfer"}
    for case in spec["cases"]:
        rows = run_e2e_fixture(ref_module, case)
        assert len(rows) == len(case["output"])
        for (node, rank), exp in zip(rows, case["output"]):
            assert node == exp["node"] and abs(rank - exp["rank"]) < spec["tolerance_abs"], case["name"]


@needs_ref_module
def test_reference_module_unit_vectors_through_host(ref_module):
    spec = json.load(open(os.path.join(GOLDEN, "pagerank_unit_vectors.json")))
    for case in spec["cases"]:
        if case["m"] != len(case["edges"]):
            continue
        gids = np.arange(case["n"], dtype=np.int64) * 5 + 11
        e = np.array(case["edges"], dtype=np.int64).reshape(-1, 2)
        with fh.Graph(gids, gids[e[:, 0]] if len(e) else [], gids[e[:, 1]] if len(e) else []) as g:
            nodes, ranks = ref_module.call(g)
        assert np.array_equal(nodes, gids)
        exp = np.array(case["expected"])
        if len(exp):
            err = np.abs(ranks - exp)
            assert err.max() < 1e-3 and err.mean() < 1e-3


@needs_ref_module
def test_host_argument_validation(ref_module):
    with fh.Graph([1, 2], [1], [2]) as g:
        with pytest.raises(fh.ProcedureError, match="must be of type FLOAT"):
            ref_module.call(g, 10, 1)  # integer literal for damping_factor is rejected (cypher_types.hpp:95,104)
        with pytest.raises(fh.ProcedureError, match="requires between 0 and 4 arguments"):
            ref_module.call(g, 10, 0.85, 1e-5, 1, 7)
        with pytest.raises(fh.ProcedureError, match=r"pagerank_reference\.get: Number of threads can't be zero \(0\)!"):
            ref_module.call(g, 10, 0.85, 1e-5, 0)


# ---- the B200 module: everything that needs no device -------------------------------------------------------

def test_b200_module_exports_and_imports():
    out = subprocess.run(["nm", "-D", "--defined-only", fh.MODULE_SO], capture_output=True, text=True, check=True).stdout
    assert " T mgp_init_module" in out and " T mgp_shutdown_module" in out
    und = subprocess.run(["nm", "-D", "--undefined-only", fh.MODULE_SO], capture_output=True, text=True, check=True).stdout
    imported = sorted(line.split()[-1] for line in und.splitlines() if " mgp_" in line)
    strong = sorted(line.split()[-1] for line in und.splitlines() if " U mgp_" in line)
    # exactly the reference module's 30 imports are strong; the optional extras are weak
    reference_imports = sorted("""mgp_list_at mgp_value_get_int mgp_value_get_double mgp_value_make_int
        mgp_value_make_double mgp_value_make_vertex mgp_value_destroy mgp_type_int mgp_type_float mgp_type_node
        mgp_module_add_read_procedure mgp_proc_add_opt_arg mgp_proc_add_result mgp_graph_approximate_vertex_count
        mgp_graph_approximate_edge_count mgp_graph_iter_vertices mgp_vertices_iterator_get mgp_vertices_iterator_next
        mgp_vertices_iterator_destroy mgp_vertex_iter_out_edges mgp_edges_iterator_get mgp_edges_iterator_next
        mgp_edges_iterator_destroy mgp_edge_get_to mgp_vertex_get_id mgp_graph_get_vertex_by_id
        mgp_graph_is_transactional mgp_result_new_record mgp_result_record_insert mgp_result_set_error_msg""".split())
    assert strong == reference_imports
    assert set(imported) - set(strong) <= {"mgp_must_abort", "mgp_result_reserve", "mgp_log"}
    # nothing CUDA-related is left undefined: the runtime is linked statically (dlopen RTLD_NOW | RTLD_LOCAL)
    assert "cuda" not in und.lower() and "nccl" not in und.lower()
    if HAVE_REF_MODULE:
        ref_und = subprocess.run(["nm", "-D", "--undefined-only", fh.REF_MODULE_SO], capture_output=True, text=True,
                                 check=True).stdout
        assert sorted(l.split()[-1] for l in ref_und.splitlines() if " U mgp_" in l) == reference_imports


def test_b200_module_signature_is_the_reference_signature(b200_module):
    assert b200_module.signature() == EXPECTED_SIGNATURE
    if HAVE_REF_MODULE:
        with fh.Module(fh.REF_MODULE_SO) as ref:
            assert ref.signature() == b200_module.signature()


def test_b200_module_error_paths_without_device(b200_module):
    before = fh.live_objects()
    with fh.Graph([5, 9, 2], [5, 9], [9, 2]) as g:
        # num_of_threads == 0: the reference's text, produced before any device work
        with pytest.raises(fh.ProcedureError, match=r"^pagerank\.get: Number of threads can't be zero \(0\)!$"):
            b200_module.call(g, 100, 0.85, 1e-5, 0)
        with pytest.raises(fh.ProcedureError, match="must be of type FLOAT"):
            b200_module.call(g, 10, 1)
        import conftest
        if not conftest.HAVE_GPU:
            # no CPU fallback: a call that needs the device fails loudly through the procedure's error message
            with pytest.raises(fh.ProcedureError, match=r"^pagerank\.get: CUDA error"):
                b200_module.call(g)
    # empty graph: zero rows, no device needed (test_empty fixture)
    with fh.Graph([], [], []) as g:
        nodes, ranks = b200_module.call(g)
        assert len(nodes) == 0 and len(ranks) == 0
        with pytest.raises(fh.ProcedureError, match="Number of threads can't be zero"):
            b200_module.call(g, 100, 0.85, 1e-5, 0)
    assert fh.live_objects() == before  # iterators / values released on the error paths too


def test_b200_module_polls_must_abort_during_ingest(b200_module):
    n = 10_000
    gids = np.arange(n, dtype=np.int64)
    with fh.Graph(gids, gids[:-1], gids[1:]) as g:
        g.set_abort(True)
        with pytest.raises(fh.ProcedureError, match="aborted"):
            b200_module.call(g)
        assert g.abort_polls() >= 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="reference checkout not present")
def test_module_compiles_against_the_reference_header(tmp_path):
    """include/mgp_abi.h restates a slice of include/mg_procedure.h; compiling the module against the
    reference header instead proves every call the module makes matches the real prototypes."""
    obj = tmp_path / "contract.o"
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-DMGB200_USE_REFERENCE_MGP_HEADER",
                    "-I", "/root/reference/include", "-I", os.path.join(REPO, "include"),
                    os.path.join(REPO, "memgraph_b200", "csrc", "pagerank_module.cpp")], check=True)
    # and the enum values / struct layout restated in mgp_abi.h agree with the reference header
    probe = tmp_path / "probe.cpp"
    probe.write_text("""
#include <cstdio>
#include <cstddef>
#include HEADER
int main() {
  printf("%d %d %d %d %zu %zu\\n", (int)mgp_error::MGP_ERROR_NO_ERROR, (int)mgp_error::MGP_ERROR_OUT_OF_RANGE,
         (int)mgp_error::MGP_ERROR_NOT_YET_IMPLEMENTED, (int)mgp_log_level::MGP_LOG_LEVEL_CRITICAL,
         sizeof(mgp_vertex_id), offsetof(mgp_vertex_id, as_int));
}
""")
    outs = []
    for header, inc in [('"mgp_abi.h"', os.path.join(REPO, "include")), ('"mg_procedure.h"', "/root/reference/include")]:
        exe = tmp_path / ("probe_" + str(len(outs)))
        subprocess.run(["g++", "-std=c++20", f"-DHEADER={header}", "-I", inc, str(probe), "-o", str(exe)], check=True)
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1] == "0 4 13 5 8 0\n"
