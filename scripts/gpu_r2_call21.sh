#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c21; O=gpurun_out/c21/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c21/err.txt | tee -a $O; }
for v in h_base h_pf8 h_pf6 h_pf5 h_pf4 h_b6 h_b4 h_u4pf8 h_u4b8; do
  export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  MGB200_TAG="$v n1 no-overlap" MGB200_OVERLAP_EPILOGUE=0 q
  MGB200_TAG="$v n1" q
  MGB200_TAG="$v lone8" MGB200_LONE_WORLD=8 q
done
