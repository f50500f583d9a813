#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c8; O=gpurun_out/c8/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c8/err.txt | tee -a $O; }
timeout 300 python -m pytest tests/test_gpu_pagerank.py -q -x 2>&1 | tail -2 | tee -a $O
for v in t_u4b6pf t_u4b8pf t_u8b4pf t_u4b6 t_u6b5pf t_u2b8pf t_hu4; do
  export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  MGB200_TAG="$v lone8" MGB200_LONE_WORLD=8 q
  MGB200_TAG="$v n1" q
done
export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/t_u4b6pf/libmgb200_pagerank.so
for w in 4 64; do
  MGB200_TAG="t_u4b6pf lone8 x$w" MGB200_SELL_WORK_ITEMS=$((148*32*w)) MGB200_LONE_WORLD=8 q
  MGB200_TAG="t_u4b6pf n1 x$w" MGB200_SELL_WORK_ITEMS=$((148*32*w)) q
done
MGB200_TAG="t_u4b6pf lone8 global" MGB200_LABELLING=global MGB200_LONE_WORLD=8 q
MGB200_TAG="t_u4b6pf lone4" MGB200_LONE_WORLD=4 q
MGB200_TAG="t_u4b6pf lone2" MGB200_LONE_WORLD=2 q
