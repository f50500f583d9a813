"""No-GPU: the C-ABI library loads and exports every function include/mgb200_*.h declares, the ctypes
binding table covers them all, and a compute call without a device fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import conftest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "memgraph_b200", "_build", "libmgb200_pagerank.so")
HEADERS = ["mgb200_pagerank.h", "mgb200_bfs.h", "mgb200_katz.h", "mgb200_personalized.h"]


def declared_functions():
    names = []
    for h in HEADERS:
        text = open(os.path.join(REPO, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments mention function names too
        names += re.findall(r"\b(mgb200_\w+)\s*\(", text)
    typedefs = {"mgb200_abort_fn"}
    return sorted(set(names) - typedefs)


def test_library_exports_every_declared_function():
    if not os.path.exists(LIB):
        from memgraph_b200 import build
        build.build_all()
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = declared_functions()
    assert len(declared) >= 22
    missing = [f for f in declared if f not in exported]
    assert not missing, missing
    # nothing but the declared API (and nothing from the oracle) leaks out of the library
    stray = sorted(s for s in exported if not s.startswith("mgb200_"))
    assert not [s for s in stray if "oracle" in s or "ref_" in s], stray
    # the CUDA runtime is linked statically: no libcudart / libnccl dependency to resolve at dlopen time
    needed = subprocess.run(["readelf", "-d", LIB], capture_output=True, text=True, check=True).stdout
    assert "libcudart" not in needed and "libnccl" not in needed


def test_ctypes_table_binds_every_declared_function():
    from memgraph_b200 import _native
    assert sorted(_native.EXPORTS) == declared_functions()
    lib = _native.lib()  # getattr on every entry: raises if a symbol is missing
    assert lib.mgb200_last_error() is not None


@pytest.mark.skipif(conftest.HAVE_GPU, reason="this is the no-device behaviour")
def test_compute_entry_points_fail_loudly_without_a_device():
    import numpy as np
    import memgraph_b200 as mg
    from memgraph_b200 import bfs
    with pytest.raises(mg.MgB200Error, match="CUDA error"):
        mg.PageRankGraph(2, 1, [[0, 1]])
    with pytest.raises(mg.MgB200Error, match="CUDA error"):
        mg.pagerank_from_edges(3, [0, 1], [1, 2])
    with pytest.raises(mg.MgB200Error, match="CUDA error"):
        mg.pagerank_from_edges(0, [], [])  # even the trivial case: the product has no CPU path
    with pytest.raises(mg.MgB200Error, match="CUDA error"):
        bfs.BfsGraph(3, [0, 1], [1, 2])
    # the reference's argument error still comes first (pagerank.cpp:63-65 precedes any work)
    with pytest.raises(mg.MgB200Error, match=r"Number of threads can't be zero \(0\)!"):
        mg.pagerank_from_edges(3, [0, 1], [1, 2], number_of_threads=0)
    s, t = mg.rmat_edges_host(5, 64)  # workload synthesis is host-side and needs no device
    assert s.max() < 32 and t.max() < 32 and len(np.unique(np.stack([s, t]), axis=1).T) > 8


def test_coo_fingerprint_is_host_only_and_order_sensitive(monkeypatch):
    """mgb200_coo_fingerprint_u32 (the key of pagerank.so's optional device-graph cache) needs no device: equal arrays give
    equal fingerprints, any changed / swapped / dropped edge a different one."""
    import ctypes
    import numpy as np
    from memgraph_b200 import _native as N
    lib = N.lib()

    def fp(n, f, t):
        f = np.ascontiguousarray(f, dtype=np.uint32)
        t = np.ascontiguousarray(t, dtype=np.uint32)
        out = (ctypes.c_uint64 * 2)()
        assert lib.mgb200_coo_fingerprint_u32(n, len(f), f.ctypes.data, t.ctypes.data, out) == 0
        return (out[0], out[1])

    rng = np.random.default_rng(1)
    for m in (0, 1, 1000, 600000):  # the last one crosses the multi-threaded threshold
        f, t = rng.integers(0, 5000, m), rng.integers(0, 5000, m)
        base = fp(5000, f, t)
        assert fp(5000, f.copy(), t.copy()) == base
        assert fp(5001, f, t) != base
        if m:
            g = f.copy(); g[m // 2] ^= 1
            assert fp(5000, g, t) != base
            assert fp(5000, f[:-1], t[:-1]) != base
        if m > 1:
            h = f.copy(); h[[0, 1]] = h[[1, 0]]
            u = t.copy(); u[[0, 1]] = u[[1, 0]]
            if (f[0], t[0]) != (f[1], t[1]):
                assert fp(5000, h, u) != base  # order-sensitive: the module's dense ids follow the iteration order
