"""The product replays std::partial_sort's tie order with index arithmetic (memgraph_b200/csrc/katz_heap.hpp -- the same
function runs on the device when a Katz convergence verdict depends on it).  Here its host instantiation is compared
with the real std::partial_sort call the reference makes (oracle/katz_oracle.cpp oracle_partial_sort_order,
katz.cpp:185-189) on inputs full of ties.  No GPU needed."""
import ctypes

import numpy as np
import pytest

from _checkers import ORACLE_SO, build_checkers


def oracle_order(keys):
    L = ctypes.CDLL(ORACLE_SO)
    if not hasattr(L, "oracle_partial_sort_order"):
        build_checkers()
        L = ctypes.CDLL(ORACLE_SO)
    L.oracle_partial_sort_order.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
    k = np.ascontiguousarray(keys, dtype=np.float64)
    out = np.zeros(len(k), dtype=np.uint32)
    assert L.oracle_partial_sort_order(len(k), k.ctypes.data, out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 16, 17, 100, 1023, 1024, 4097])
@pytest.mark.parametrize("distinct", [1, 2, 3, 10, 10**9])
def test_tie_order_equals_std_partial_sort(n, distinct):
    from memgraph_b200 import katz
    rng = np.random.default_rng(n * 31 + distinct % 1000)
    for trial in range(3):
        keys = rng.integers(0, distinct, n).astype(np.float64) * 0.25
        got = katz.tie_order(keys)
        ref = oracle_order(keys)
        assert np.array_equal(got, ref), (n, distinct, trial)
        assert np.all(np.diff(keys[got]) <= 0)  # and it is sorted by key descending


def test_tie_order_on_sorted_and_reversed_input():
    from memgraph_b200 import katz
    for keys in (np.arange(50, dtype=np.float64), np.arange(50, dtype=np.float64)[::-1].copy(), np.zeros(50),
                 np.repeat(np.arange(5.0), 10), np.tile(np.arange(5.0), 10)):
        assert np.array_equal(katz.tie_order(keys), oracle_order(keys))


def verdict_like_the_kernels(c, ur, eps):
    """numpy restatement of katz.cu's group_mark / member_test / group_verdict / tie_check kernels."""
    from memgraph_b200 import katz
    n = len(c)
    rows = np.argsort(-c, kind="stable")
    ks = c[rows]
    start = np.zeros(n, dtype=np.int64)
    for i in range(1, n):
        start[i] = i if ks[i - 1] != ks[i] else start[i - 1]
    bound = ur[rows] - eps
    definite_fail = bool(np.any((start > 0) & (bound >= ks[np.maximum(start - 1, 0)])))
    strict = bound >= ks
    ambiguous = []
    for s in np.unique(start):
        members = np.flatnonzero(start == s)
        cnt = int(strict[members].sum())
        if cnt >= 2:
            definite_fail = True
        if cnt == 1 and len(members) >= 2:
            ambiguous.append((s, rows[members[strict[members]][0]]))
    if definite_fail:
        return False, len(ambiguous)
    if ambiguous:
        order = katz.tie_order(c)
        for s, row in ambiguous:
            if order[s] != row:
                return False, len(ambiguous)
    return True, len(ambiguous)


def verdict_like_the_reference(c, ur, eps):
    order = oracle_order(c)  # Converged(), katz.cpp:185-209
    for i in range(1, len(c)):
        if ur[order[i]] - eps >= c[order[i - 1]]:
            return False
    return True


def test_group_analysis_gives_the_references_verdict():
    """The verdict depends on the order inside groups of equal centralities only for groups with exactly one
    self-violating member -- and there the replayed partial_sort order settles it.  Brute force over small cases,
    including many where that rare case occurs."""
    rng = np.random.default_rng(11)
    seen_ambiguous = seen_true = seen_false = 0
    for trial in range(4000):
        n = int(rng.integers(1, 12))
        c = rng.integers(0, 4, n).astype(np.float64)
        ur = c + rng.choice([0.0, 0.5, 1.0, 2.0], n)
        eps = float(rng.choice([0.25, 0.75, 1.5]))
        got, amb = verdict_like_the_kernels(c, ur, eps)
        ref = verdict_like_the_reference(c, ur, eps)
        assert got == ref, (c, ur, eps)
        seen_ambiguous += amb > 0
        seen_true += ref
        seen_false += not ref
    assert seen_ambiguous > 50 and seen_true > 50 and seen_false > 50
