"""ctypes front-end of memgraph_b200/_build/libmgp_fake_host.so (the in-memory mgp host, test
infrastructure).  Loaded RTLD_GLOBAL so that query modules dlopen'ed through it bind their
undefined mgp_* symbols to it, like they bind to the memgraph executable."""
import ctypes
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(REPO, "memgraph_b200", "_build", "libmgp_fake_host.so")
MODULE_SO = os.path.join(REPO, "memgraph_b200", "_build", "pagerank.so")
REF_MODULE_SO = os.path.join(REPO, "oracle", "_ref", "pagerank_reference.so")

_host = None


def host():
    global _host
    if _host is None:
        if not os.path.exists(HOST_SO):
            from memgraph_b200 import build
            build.build_all()
        H = ctypes.CDLL(HOST_SO, mode=ctypes.RTLD_GLOBAL)
        vp, i64p, f64p = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        H.fh_graph_create.restype = vp
        H.fh_graph_create.argtypes = [ctypes.c_uint64, i64p, ctypes.c_uint64, i64p, i64p, ctypes.c_int]
        H.fh_graph_destroy.argtypes = [vp]
        H.fh_graph_set_abort.argtypes = [vp, ctypes.c_int]
        H.fh_graph_abort_polls.restype = ctypes.c_long
        H.fh_graph_abort_polls.argtypes = [vp]
        H.fh_graph_hide_vertex.argtypes = [vp, ctypes.c_int64]
        H.fh_live_objects.restype = ctypes.c_long
        H.fh_module_load.restype = vp
        H.fh_module_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        H.fh_module_close.argtypes = [vp]
        H.fh_module_signature.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        H.fh_call.restype = vp
        H.fh_call.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p]
        H.fh_call_ex.restype = vp
        H.fh_call_ex.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        H.fh_graph_set_edge_property.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint64, i64p, f64p, ctypes.c_int]
        H.fh_result_error.restype = ctypes.c_char_p
        H.fh_result_error.argtypes = [vp]
        H.fh_result_rows.restype = ctypes.c_uint64
        H.fh_result_rows.argtypes = [vp]
        H.fh_result_copy.argtypes = [vp, ctypes.c_void_p, ctypes.c_void_p]
        H.fh_result_destroy.argtypes = [vp]
        _host = H
    return _host


BFS_MODULE_SO = os.path.join(REPO, "memgraph_b200", "_build", "gpu_bfs.so")
CUGRAPH_PAGERANK_SO = os.path.join(REPO, "memgraph_b200", "_build", "cugraph.pagerank.so")
CUGRAPH_PERSONALIZED_SO = os.path.join(REPO, "memgraph_b200", "_build", "cugraph.personalized_pagerank.so")


class Node:
    """A NODE argument, passed to the procedure by the vertex's gid."""

    def __init__(self, gid):
        self.gid = int(gid)


class ProcedureError(RuntimeError):
    """What the engine raises as QueryRuntimeException("<module>.<proc>: <msg>")."""


class Graph:
    def __init__(self, gids, src_gid, dst_gid, transactional=True):
        g = np.ascontiguousarray(gids, dtype=np.int64)
        s = np.ascontiguousarray(src_gid, dtype=np.int64)
        d = np.ascontiguousarray(dst_gid, dtype=np.int64)
        self.h = host().fh_graph_create(len(g), g.ctypes.data, len(s), s.ctypes.data, d.ctypes.data,
                                        1 if transactional else 0)
        if not self.h:
            raise ValueError("edge endpoint is not a vertex of the graph")

    def set_edge_property(self, name, src_gid, values, as_int=False):
        """One numeric property on the edges, in the order they were given to the constructor (NaN = edge lacks it)."""
        s = np.ascontiguousarray(src_gid, dtype=np.int64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        if host().fh_graph_set_edge_property(self.h, name.encode(), len(v), s.ctypes.data, v.ctypes.data, 1 if as_int else 0):
            raise ValueError("edge property does not match the graph's edges")

    def set_abort(self, flag=True):
        host().fh_graph_set_abort(self.h, 1 if flag else 0)

    def abort_polls(self):
        return host().fh_graph_abort_polls(self.h)

    def hide_vertex(self, gid):
        assert host().fh_graph_hide_vertex(self.h, gid) == 0

    def close(self):
        if self.h:
            host().fh_graph_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Module:
    """A loaded query module (SharedLibraryModule)."""

    def __init__(self, path):
        err = ctypes.create_string_buffer(1024)
        self.h = host().fh_module_load(path.encode(), err, 1024)
        if not self.h:
            raise OSError(err.value.decode())

    def signature(self, proc="get"):
        buf = ctypes.create_string_buffer(1024)
        if host().fh_module_signature(self.h, proc.encode(), buf, 1024):
            raise KeyError(proc)
        return buf.value.decode()

    def call(self, graph, *args, proc="get"):
        """args: Python ints are INTEGER literals, floats are FLOAT literals (strictly typed).
        Returns (node_gids, ranks) in emission order."""
        def kind(a):
            if isinstance(a, Node):
                return "v"
            if isinstance(a, str):
                return "s"
            if isinstance(a, (list, tuple)):  # LIST OF NODE / LIST OF FLOAT (an empty list is typed by the signature slot)
                return "V" if (a and isinstance(a[0], Node)) else "D"
            return "i" if isinstance(a, (int, np.integer)) and not isinstance(a, bool) else "d"

        kinds = "".join(kind(a) for a in args)
        if "D" in kinds:  # an empty Python list carries no element type: take it from the procedure's signature
            sig = self.signature(proc)
            slots = sig[sig.index("(") + 1:sig.index(") ::")].split(", ")
            kinds = "".join("V" if (k == "D" and not a and "LIST OF NODE" in slots[i]) else k
                            for i, (k, a) in enumerate(zip(kinds, args)))
        iv = np.array([a.gid if k == "v" else (int(a) if k == "i" else 0) for a, k in zip(args, kinds)] + [0],
                      dtype=np.int64)
        dv = np.array([float(a) if k == "d" else 0.0 for a, k in zip(args, kinds)] + [0.0], dtype=np.float64)
        sv = (ctypes.c_char_p * (len(args) + 1))(*[a.encode() if k == "s" else None for a, k in zip(args, kinds)], None)
        ll = np.array([len(a) if k in "VD" else 0 for a, k in zip(args, kinds)] + [0], dtype=np.uint64)
        li = np.array([x.gid for a, k in zip(args, kinds) if k == "V" for x in a] + [0], dtype=np.int64)
        ld = np.array([float(x) for a, k in zip(args, kinds) if k == "D" for x in a] + [0.0], dtype=np.float64)
        # fh_call_ex walks list_ivals / list_dvals with ONE cursor over all list arguments in order: interleave accordingly
        if "V" in kinds and "D" in kinds:
            li_full, ld_full = [], []
            for a, k in zip(args, kinds):
                if k == "V":
                    li_full += [x.gid for x in a]; ld_full += [0.0] * len(a)
                elif k == "D":
                    li_full += [0] * len(a); ld_full += [float(x) for x in a]
            li = np.array(li_full + [0], dtype=np.int64)
            ld = np.array(ld_full + [0.0], dtype=np.float64)
        r = host().fh_call_ex(self.h, proc.encode(), graph.h, len(args), kinds.encode(), iv.ctypes.data, dv.ctypes.data,
                              sv, ll.ctypes.data, li.ctypes.data, ld.ctypes.data)
        try:
            err = host().fh_result_error(r)
            if err is not None:
                raise ProcedureError(err.decode())
            n = host().fh_result_rows(r)
            nodes = np.zeros(n, dtype=np.int64)
            ranks = np.zeros(n, dtype=np.float64)
            host().fh_result_copy(r, nodes.ctypes.data, ranks.ctypes.data)
            return nodes, ranks
        finally:
            host().fh_result_destroy(r)

    def close(self):
        if self.h:
            rc = host().fh_module_close(self.h)
            self.h = None
            return rc

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def live_objects():
    return host().fh_live_objects()
