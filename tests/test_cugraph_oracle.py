"""oracle/cugraph_pagerank_oracle.c (the checker of the personalised / weighted variant) against the
reference's own e2e fixtures for its cuGraph modules (tests/golden/cugraph_pagerank_e2e.json, extracted by
tests/golden/make_cugraph_golden.py from mage/tests/e2e/*/test_cugraph_*).  cuGraph itself is a third-party library
that is absent here; these fixtures (3-6 printed digits) are what pins the restatement."""
import ctypes
import json
import os

import numpy as np
import pytest

from _checkers import ORACLE_SO, build_checkers

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = json.load(open(os.path.join(HERE, "golden", "cugraph_pagerank_e2e.json")))


def oracle_cugraph_pagerank(n, f, t, weights=None, personalization=None, alpha=0.85, epsilon=1e-5, max_iterations=100):
    """-> (ranks, iterations, converged); personalization = (dense vertex ids, values) or None."""
    if not os.path.exists(ORACLE_SO):
        build_checkers()
    L = ctypes.CDLL(ORACLE_SO)
    L.oracle_cugraph_pagerank.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_uint64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64),
                                          ctypes.POINTER(ctypes.c_int)]
    f = np.ascontiguousarray(f, dtype=np.uint64)
    t = np.ascontiguousarray(t, dtype=np.uint64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    pv = pw = None
    if personalization is not None:
        pv = np.ascontiguousarray(personalization[0], dtype=np.uint64)
        pw = np.ascontiguousarray(personalization[1], dtype=np.float64)
    out = np.zeros(n)
    it, conv = ctypes.c_uint64(0), ctypes.c_int(0)
    rc = L.oracle_cugraph_pagerank(n, len(f), f.ctypes.data, t.ctypes.data, None if w is None else w.ctypes.data,
                                   0 if pv is None else len(pv), None if pv is None else pv.ctypes.data,
                                   None if pw is None else pw.ctypes.data, alpha, epsilon, max_iterations,
                                   out.ctypes.data, ctypes.byref(it), ctypes.byref(conv))
    if rc:
        raise ValueError(f"oracle_cugraph_pagerank: {rc}")
    return out, it.value, bool(conv.value)


def fixture_arrays(fx):
    dense = {g: i for i, g in enumerate(fx["nodes"])}
    f = [dense[a] for a, _ in fx["edges"]]
    t = [dense[b] for _, b in fx["edges"]]
    pers = None
    if fx["personalization"]:
        pers = ([dense[g] for g, _ in fx["personalization"]], [v for _, v in fx["personalization"]])
    return dense, f, t, pers


@pytest.mark.parametrize("fx", [f for f in FIXTURES if f["nodes"]], ids=lambda f: f["name"])
def test_reference_e2e_fixtures(fx):
    dense, f, t, pers = fixture_arrays(fx)
    ranks, it, conv = oracle_cugraph_pagerank(len(dense), f, t, personalization=pers)  # module defaults 100, 0.85, 1e-5
    assert conv
    for gid, want in fx["expected"]:
        digits = len(str(want).split(".")[1]) if "." in str(want) else 0
        tol = max(1.5 * 10.0 ** -digits, 2e-5)  # printed digits; the solver itself stops at an L1 residual of 1e-5
        assert abs(ranks[dense[gid]] - want) < tol, (fx["name"], gid, ranks[dense[gid]], want)


def test_restatement_properties():
    rng = np.random.default_rng(3)
    n, m = 300, 2000
    f, t = rng.integers(0, n, m), rng.integers(0, n, m)
    base, it, conv = oracle_cugraph_pagerank(n, f, t, epsilon=1e-12, max_iterations=500)
    assert conv and abs(base.sum() - 1.0) < 1e-9          # dangling mass is redistributed: the vector stays stochastic
    uni, _, _ = oracle_cugraph_pagerank(n, f, t, personalization=(np.arange(n), np.full(n, 3.0)), epsilon=1e-12,
                                        max_iterations=500)
    assert np.max(np.abs(uni - base)) < 1e-12             # uniform personalisation == none (values are normalised)
    w1, _, _ = oracle_cugraph_pagerank(n, f, t, weights=np.full(m, 2.5), epsilon=1e-12, max_iterations=500)
    assert np.max(np.abs(w1 - base)) < 1e-12              # a constant weight cancels against the out-weight sums
    capped, it, conv = oracle_cugraph_pagerank(n, f, t, epsilon=0.0, max_iterations=7)
    assert it == 7 and not conv
