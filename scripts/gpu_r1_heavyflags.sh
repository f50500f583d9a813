#!/bin/bash
# heavy_segments_kernel<kPathFlags> variants on one B200 (MGB200_IDX_FLAGS=1): register cap and per-load L2 selection
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/heavyflags.txt; : > $O
V=memgraph_b200/_build/variants
MGB200_TAG="range default" timeout 40 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
for v in hv_base hv_b6 hv_l2u hv_l2u_b6; do
  MGB200_TAG="flags $v" MGB200_IDX_FLAGS=1 MGB200_LIBRARY=$V/$v/libmgb200_pagerank.so timeout 40 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
done
