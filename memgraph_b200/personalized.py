"""cuGraph-semantics PageRank on the device (include/mgb200_personalized.h): the host-side mirror of what
``cugraph.pagerank.get`` / ``cugraph.personalized_pagerank.get`` compute (mage/cpp/cugraph_module/algorithms/pagerank.cu,
personalized_pagerank.cu).  Runs on a :class:`memgraph_b200.pagerank.PageRankGraph` handle."""
import ctypes

import numpy as np

from . import _native as N
from .pagerank import PageRankGraph, _check


def cugraph_pagerank(graph, personalization_vertices=None, personalization_values=None, max_iterations=100,
                     damping_factor=0.85, stop_epsilon=1e-5):
    """-> (pageranks[n] in original vertex-id order, stats dict).  Seeds are dense vertex ids; no seeds = plain PageRank."""
    n = graph.get_node_count()
    p = N.CugraphParams()
    p.max_iterations, p.damping_factor, p.stop_epsilon = int(max_iterations), float(damping_factor), float(stop_epsilon)
    keep = []
    if personalization_vertices is not None and len(personalization_vertices):
        v = np.ascontiguousarray(personalization_vertices, dtype=np.uint64)
        w = np.ascontiguousarray(personalization_values, dtype=np.float64)
        if len(v) != len(w):
            raise ValueError("personalization_vertices and personalization_values must have the same length.")
        keep = [v, w]
        p.n_personalization, p.personalization_vertices, p.personalization_values = len(v), v.ctypes.data, w.ctypes.data
    out = np.zeros(n, dtype=np.float64)
    st = N.CugraphStats()
    _check(N.lib().mgb200_cugraph_pagerank_run(graph.handle, ctypes.byref(p), out.ctypes.data if n else None,
                                               ctypes.byref(st)))
    del keep
    return out, {f: getattr(st, f) for f, _ in N.CugraphStats._fields_}


def weighted_graph(n, sources, targets, weights, device=0):
    """A single-partition handle with one FP64 weight per edge (mgb200_graph_create_host_weighted_u32)."""
    f = np.ascontiguousarray(sources, dtype=np.uint32)
    t = np.ascontiguousarray(targets, dtype=np.uint32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if not (len(f) == len(t) == len(w)):
        raise ValueError("sources, targets and weights must have the same length")
    g = PageRankGraph.__new__(PageRankGraph)
    h = N.vp()
    _check(N.lib().mgb200_graph_create_host_weighted_u32(device, int(n), len(f), f.ctypes.data, t.ctypes.data, w.ctypes.data,
                                                         ctypes.byref(h)))
    g._finish(h, device)
    return g


def cugraph_pagerank_from_edges(n, sources, targets, device=0, weights=None, **kw):
    g = (PageRankGraph.from_arrays(n, sources, targets, device=device) if weights is None
         else weighted_graph(n, sources, targets, weights, device=device))
    with g:
        return cugraph_pagerank(g, **kw)
