"""Static Katz centrality on the device (include/mgb200_katz.h): the host-side mirror of
``katz_alg::SetKatz(graph, alpha = 0.2, epsilon = 1e-2)`` (mage/cpp/katz_centrality_module/algorithm/katz.hpp:34-35).
Runs on a :class:`memgraph_b200.pagerank.PageRankGraph` handle -- the two paths share one device-resident layout."""
import ctypes

import numpy as np

from . import _native as N
from .pagerank import PageRankGraph, _check


class NotConverged(RuntimeError):
    """max_iterations (a caller-side guard; the reference has none) ended the loop; ``.centralities`` holds the
    last iterate."""

    def __init__(self, centralities, stats):
        super().__init__("Katz centrality: max_iterations reached before the ranking separated")
        self.centralities, self.stats = centralities, stats


def set_katz(graph, alpha=0.2, epsilon=1e-2, max_iterations=0):
    """-> (centralities[n] in original vertex-id order, stats dict)."""
    n = graph.get_node_count()
    out = np.zeros(n, dtype=np.float64)
    st = N.KatzStats()
    rc = N.lib().mgb200_katz_run(graph.handle, float(alpha), float(epsilon), int(max_iterations), out.ctypes.data,
                                 ctypes.byref(st))
    stats = {f: getattr(st, f) for f, _ in N.KatzStats._fields_}
    if rc == N.KATZ_NOT_CONVERGED:
        raise NotConverged(out, stats)
    _check(rc)
    return out, stats


def katz_from_edges(n, sources, targets, alpha=0.2, epsilon=1e-2, max_iterations=0, device=0):
    with PageRankGraph.from_arrays(n, sources, targets, device=device) as g:
        return set_katz(g, alpha, epsilon, max_iterations)


def tie_order(keys):
    """The order std::partial_sort(first, last, last, key-descending) leaves ids 0..n-1 in (host-only diagnostic)."""
    k = np.ascontiguousarray(keys, dtype=np.float64)
    order = np.zeros(len(k), dtype=np.uint32)
    _check(N.lib().mgb200_katz_tie_order(len(k), k.ctypes.data, order.ctypes.data))
    return order
