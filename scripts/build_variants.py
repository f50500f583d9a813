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
    "hv_base": [],
    "hv_b6": ["-DMGB_HEAVY_MIN_BLOCKS=6"],
    "hv_l2u": ["-DMGB_HEAVY_FLAGS_L2=0"],
    "hv_l2u_b6": ["-DMGB_HEAVY_FLAGS_L2=0", "-DMGB_HEAVY_MIN_BLOCKS=6"],
    "s_w8_i3_v4": [],
    "s_w8_i3_v6": ["-DMGB_STREAM_VALS=6"],
    "s_w8_i2_v8": ["-DMGB_STREAM_IDX_STAGES=2", "-DMGB_STREAM_VALS=8"],
    "s_w12_i2_v4": ["-DMGB_STREAM_WARPS=12", "-DMGB_STREAM_IDX_STAGES=2", "-DMGB_STREAM_VALS=4"],
    "s_w16_i2_v3": ["-DMGB_STREAM_WARPS=16", "-DMGB_STREAM_IDX_STAGES=2", "-DMGB_STREAM_VALS=3"],
    "s_w4_i4_v8": ["-DMGB_STREAM_WARPS=4", "-DMGB_STREAM_IDX_STAGES=4", "-DMGB_STREAM_VALS=8"],
    "s_w8_i3_v2": ["-DMGB_STREAM_VALS=2"],
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
        regs = [l for l in log.stdout.splitlines() if "sell_rows" in l or ("Used" in l)]
        sell = ""
        lines = log.stdout.splitlines()
        for i, l in enumerate(lines):
            if "Compiling entry function" in l and "sell_stream" in l:
                sell = " | ".join(x.strip() for x in lines[i + 1:i + 4])
        subprocess.run([B.NVCC, "-shared", "-cudart", "static", "-o", os.path.join(out, "libmgb200_pagerank.so"), obj] + others, check=True)
        print(f"{name:10s} {sell}")


if __name__ == "__main__":
    main()
