#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/sweep1.txt; : > $O
V=memgraph_b200/_build/variants
run() { # tag, env...
  tag=$1; shift
  env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
}
# parity first on the default build
timeout 900 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "not scale26 and not scale22" 2>&1 | tail -3 | tee -a $O
for mb in 0 32 48 56 64 80 128; do run "default hot=$mb" MGB200_L2_HOT_MB=$mb; done
for v in base pf pf_b5 pf_b6 b6 b8 u4_pf_b8 u4_b8 u16_b3; do run "variant=$v" MGB200_LIBRARY=$V/$v/libmgb200_pagerank.so; done
for h in 64 256 4096 16384; do run "heavy_min=$h" MGB200_HEAVY_MIN_DEGREE=$h; done
for sg in 1024 16384; do run "seg=$sg" MGB200_SEGMENT_EDGES=$sg; done
cat $O
