#!/usr/bin/env python3
"""Builds kernel variants (different -D tunables) of libmgb200_pagerank.so for on-GPU sweeps.
Output: memgraph_b200/_build/variants/<name>/libmgb200_pagerank.so (select with MGB200_LIBRARY)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from memgraph_b200 import build as B  # noqa: E402

VARIANTS = {
    "h_base": [], "h_pf8": ["-DMGB_HEAVY_PREFETCH=1"], "h_pf6": ["-DMGB_HEAVY_PREFETCH=1", "-DMGB_HEAVY_MIN_BLOCKS=6"],
    "h_pf5": ["-DMGB_HEAVY_PREFETCH=1", "-DMGB_HEAVY_MIN_BLOCKS=5"], "h_pf4": ["-DMGB_HEAVY_PREFETCH=1", "-DMGB_HEAVY_MIN_BLOCKS=4"],
    "h_b6": ["-DMGB_HEAVY_MIN_BLOCKS=6"], "h_b4": ["-DMGB_HEAVY_MIN_BLOCKS=4"],
    "h_u4pf8": ["-DMGB_HEAVY_PREFETCH=1", "-DMGB_UNROLL=4"], "h_u4b8": ["-DMGB_UNROLL=4"],
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    objs = B.build_core()
    others = [o for o in objs if not o.endswith("pagerank_kernels.o")]
    for name in names:
        out = os.path.join(B.OUT, "variants", name)
        os.makedirs(out, exist_ok=True)
        obj = os.path.join(out, "pagerank_kernels.o")
        log = subprocess.run([B.NVCC] + B.NVCC_FLAGS + VARIANTS[name] + ["-c", os.path.join(B.CSRC, "pagerank_kernels.cu"), "-o", obj],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if log.returncode:
            print(log.stdout)
            raise SystemExit(1)
        sell = ""
        lines = log.stdout.splitlines()
        for i, l in enumerate(lines):
            if "Compiling entry function" in l and "heavy_segments_kernelILi0ELb0E" in l:
                sell = " | ".join(x.strip() for x in lines[i + 1:i + 4])
        subprocess.run([B.NVCC, "-shared", "-cudart", "static", "-o", os.path.join(out, "libmgb200_pagerank.so"), obj] + others, check=True)
        print(f"{name:10s} {sell}")


if __name__ == "__main__":
    main()
