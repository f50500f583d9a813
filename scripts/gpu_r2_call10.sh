#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c10; O=gpurun_out/c10/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c10/err.txt | tee -a $O; }
timeout 300 python -m pytest tests/test_gpu_pagerank.py tests/test_gpu_katz.py -q -x 2>&1 | tail -2 | tee -a $O
for v in s_u8b4 s_u4b6 s_u6b5 s_u4b8 s_u8b5; do
  export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  for w in 1 4 16; do
    MGB200_TAG="$v lone8 ticket x$w" MGB200_SELL_WORK_ITEMS=$((148*32*w)) MGB200_LONE_WORLD=8 q
    MGB200_TAG="$v lone8 static x$w" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32*w)) MGB200_LONE_WORLD=8 q
  done
  MGB200_TAG="$v n1 ticket x16" q
  MGB200_TAG="$v n1 ticket x4" MGB200_SELL_WORK_ITEMS=$((148*32*4)) q
  MGB200_TAG="$v n1 static x1" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32)) q
done
