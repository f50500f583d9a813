"""In-tree build of the native artefacts (nvcc cross-compiles sm_100a without a GPU).

    memgraph_b200/_build/libmgb200_pagerank.so   C ABI of include/mgb200_pagerank.h (CUDA runtime linked statically)
    memgraph_b200/_build/pagerank.so             drop-in Memgraph query module (mgp_init_module / mgp_shutdown_module)
    memgraph_b200/_build/libmgp_fake_host.so     fake mgp host used by the tests to load query modules without Memgraph

Objects are rebuilt only when a source or header is newer.  `python -m memgraph_b200.build` builds everything.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "_build")
INCLUDE = os.path.join(REPO, "include")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-I", INCLUDE, "-I", CSRC,
    "-diag-suppress", "1444",  # cub::TransformInputIterator deprecation notice
]
CUDA_SOURCES = ["pagerank_kernels.cu", "graph_build.cu", "capi.cu", "bfs.cu", "katz.cu", "personalized.cu"]
HEADERS = [os.path.join(CSRC, "core.hpp"), os.path.join(CSRC, "rmat.hpp"), os.path.join(CSRC, "sell_stream.cuh"),
           os.path.join(INCLUDE, "mgb200_pagerank.h"), os.path.join(INCLUDE, "mgb200_bfs.h"),
           os.path.join(INCLUDE, "mgb200_katz.h"), os.path.join(INCLUDE, "mgb200_personalized.h"), os.path.join(CSRC, "katz_heap.hpp"),
           os.path.join(INCLUDE, "mgp_abi.h"), os.path.join(CSRC, "mgp_module_common.hpp")]

CORE_LIB = os.path.join(OUT, "libmgb200_pagerank.so")
MODULE_LIB = os.path.join(OUT, "pagerank.so")
FAKE_HOST_LIB = os.path.join(OUT, "libmgp_fake_host.so")
BFS_MODULE_LIB = os.path.join(OUT, "gpu_bfs.so")
KATZ_MODULE_LIB = os.path.join(OUT, "katz_centrality.so")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd, log_name=None):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log_name:
        with open(os.path.join(OUT, log_name), "w") as f:
            f.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return proc.stdout


def build_core(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for src in CUDA_SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        if _newer(obj, [path] + HEADERS):
            if verbose:
                print("nvcc", src, flush=True)
            _run([NVCC] + NVCC_FLAGS + ["-c", path, "-o", obj], log_name=src + ".ptxas.log")
        objs.append(obj)
    if _newer(CORE_LIB, objs):
        # -fvisibility=hidden + explicit default visibility on the extern "C" API keeps the export list tight
        _run([NVCC, "-shared", "-cudart", "static", "-o", CORE_LIB] + objs + ["-Xlinker", "--no-undefined"])
    return objs


def build_module(objs, verbose=False):
    """The drop-in query module: module glue + the same CUDA objects, CUDA runtime static, so that
    dlopen(RTLD_NOW | RTLD_LOCAL) needs nothing but the host's mgp_* symbols (module.cpp:861)."""
    src = os.path.join(CSRC, "pagerank_module.cpp")
    obj = os.path.join(OUT, "pagerank_module.o")
    if _newer(obj, [src] + HEADERS):
        if verbose:
            print("g++ pagerank_module.cpp", flush=True)
        _run([CXX, "-std=c++20", "-O2", "-fPIC", "-fvisibility=hidden", "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj])
    if _newer(MODULE_LIB, objs + [obj]):
        _run([NVCC, "-shared", "-cudart", "static", "-o", MODULE_LIB, obj] + objs)


def build_bfs_module(objs, verbose=False):
    """gpu_bfs.so: the breadth-first expansion as a read procedure (same linking rules as pagerank.so)."""
    src = os.path.join(CSRC, "gpu_bfs_module.cpp")
    obj = os.path.join(OUT, "gpu_bfs_module.o")
    if _newer(obj, [src] + HEADERS):
        if verbose:
            print("g++ gpu_bfs_module.cpp", flush=True)
        _run([CXX, "-std=c++20", "-O2", "-fPIC", "-fvisibility=hidden", "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj])
    if _newer(BFS_MODULE_LIB, objs + [obj]):
        _run([NVCC, "-shared", "-cudart", "static", "-o", BFS_MODULE_LIB, obj] + objs)


def build_katz_module(objs, verbose=False):
    """katz_centrality.so: static Katz centrality behind the reference module's name and signature."""
    src = os.path.join(CSRC, "katz_centrality_module.cpp")
    obj = os.path.join(OUT, "katz_centrality_module.o")
    if _newer(obj, [src] + HEADERS):
        if verbose:
            print("g++ katz_centrality_module.cpp", flush=True)
        _run([CXX, "-std=c++20", "-O2", "-fPIC", "-fvisibility=hidden", "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj])
    if _newer(KATZ_MODULE_LIB, objs + [obj]):
        _run([NVCC, "-shared", "-cudart", "static", "-o", KATZ_MODULE_LIB, obj] + objs)


CUGRAPH_MODULE_LIBS = {0: os.path.join(OUT, "cugraph.pagerank.so"), 1: os.path.join(OUT, "cugraph.personalized_pagerank.so")}


def build_cugraph_modules(objs, verbose=False):
    """cugraph.pagerank.so / cugraph.personalized_pagerank.so: stand-ins for the reference's cuGraph PageRank modules
    (one source, compiled twice; module name = file stem, so the procedures are cugraph.pagerank.get and
    cugraph.personalized_pagerank.get)."""
    src = os.path.join(CSRC, "cugraph_modules.cpp")
    for personalized, lib in CUGRAPH_MODULE_LIBS.items():
        obj = os.path.join(OUT, f"cugraph_module_{personalized}.o")
        if _newer(obj, [src] + HEADERS):
            if verbose:
                print(f"g++ cugraph_modules.cpp (personalized={personalized})", flush=True)
            _run([CXX, "-std=c++20", "-O2", "-fPIC", "-fvisibility=hidden", f"-DMGB200_CUGRAPH_PERSONALIZED={personalized}",
                  "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj])
        if _newer(lib, objs + [obj]):
            _run([NVCC, "-shared", "-cudart", "static", "-o", lib, obj] + objs)


FAKE_HOST_SRC = os.path.join(REPO, "tests", "mgp_fake_host.cpp")  # test infrastructure: lives with the tests


def build_fake_host(verbose=False):
    src = FAKE_HOST_SRC
    if _newer(FAKE_HOST_LIB, [src] + HEADERS):
        if verbose:
            print("g++ mgp_fake_host.cpp", flush=True)
        _run([CXX, "-std=c++20", "-O2", "-fPIC", "-shared", "-I", INCLUDE, "-o", FAKE_HOST_LIB, src, "-ldl"])


def build_all(verbose=False):
    objs = build_core(verbose)
    if os.path.exists(os.path.join(CSRC, "pagerank_module.cpp")):
        build_module(objs, verbose)
    if os.path.exists(FAKE_HOST_SRC):
        build_fake_host(verbose)
    if os.path.exists(os.path.join(CSRC, "gpu_bfs_module.cpp")):
        build_bfs_module(objs, verbose)
    out = {"core": CORE_LIB}
    if os.path.exists(os.path.join(CSRC, "cugraph_modules.cpp")):
        build_cugraph_modules(objs, verbose)
        out["cugraph_pagerank_module"] = CUGRAPH_MODULE_LIBS[0]
        out["cugraph_personalized_pagerank_module"] = CUGRAPH_MODULE_LIBS[1]
    if os.path.exists(os.path.join(CSRC, "katz_centrality_module.cpp")):
        build_katz_module(objs, verbose)
        out["katz_module"] = KATZ_MODULE_LIB
    if os.path.exists(os.path.join(CSRC, "gpu_bfs_module.cpp")):
        out["bfs_module"] = BFS_MODULE_LIB
    if os.path.exists(os.path.join(CSRC, "pagerank_module.cpp")):
        out["module"] = MODULE_LIB
    if os.path.exists(FAKE_HOST_SRC):
        out["fake_host"] = FAKE_HOST_LIB
    return out


if __name__ == "__main__":
    print(build_all(verbose=True))
