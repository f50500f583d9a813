#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/sweep3.txt; : > $O
V=memgraph_b200/_build/variants
run() { tag=$1; shift; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O; }
for v in base exp1_noepi exp2_hotgather exp3_noidx exp4_hot_noidx exp5_all; do run "variant=$v" MGB200_LIBRARY=$V/$v/libmgb200_pagerank.so; done
