"""Host-side mirror of the breadth-first expansion (include/mgb200_bfs.h): the arithmetic of
`MATCH (a)-[*BFS lower..upper]->(b)` (SingleSourceShortestPathCursor, src/query/plan/operator.cpp:2692-2912)."""
import ctypes

import numpy as np

from . import _native as N
from .pagerank import _check

OUT, IN, BOTH = 0, 1, 2  # EdgeAtom::Direction
INT64_MAX = 2**63 - 1


class BfsGraph:
    """Device-resident CSR + CSC for traversals."""

    def __init__(self, number_of_nodes, sources, targets, device=0):
        s = np.ascontiguousarray(sources, dtype=np.uint64)
        t = np.ascontiguousarray(targets, dtype=np.uint64)
        h = N.vp()
        _check(N.lib().mgb200_bfs_graph_create_host(device, int(number_of_nodes), len(s), s.ctypes.data, t.ctypes.data,
                                                    ctypes.byref(h)))
        self._h, self.n, self.m = h, int(number_of_nodes), len(s)

    @classmethod
    def from_device(cls, number_of_nodes, number_of_edges, d_sources_ptr, d_targets_ptr, device=0):
        self = cls.__new__(cls)
        h = N.vp()
        _check(N.lib().mgb200_bfs_graph_create_device(device, int(number_of_nodes), int(number_of_edges), d_sources_ptr,
                                                      d_targets_ptr, ctypes.byref(h)))
        self._h, self.n, self.m = h, int(number_of_nodes), int(number_of_edges)
        return self

    def distances(self, source, direction=OUT, lower_bound=1, upper_bound=INT64_MAX):
        """depth at which the reference emits each vertex, -1 where it does not; returns (int32[n], stats dict)."""
        out = np.empty(self.n, dtype=np.int32)
        st = N.BfsStats()
        _check(N.lib().mgb200_bfs_run(self._h, int(source), int(direction), int(lower_bound), int(upper_bound),
                                      out.ctypes.data if self.n else None, 0, ctypes.byref(st)))
        return out, {k: getattr(st, k) for k, _ in N.BfsStats._fields_}

    def close(self):
        if getattr(self, "_h", None):
            N.lib().mgb200_bfs_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
