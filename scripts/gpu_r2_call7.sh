#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c7; O=gpurun_out/c7/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c7/err.txt | tee -a $O; }
for v in base pf u4b8 u4b6 u4b6pf u16b2 u8b5 u8b5pf; do
  export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  MGB200_TAG="$v lone8 static x1" MGB200_SELL_SCHED=static MGB200_SELL_WORK_ITEMS=$((148*32)) MGB200_LONE_WORLD=8 q
  MGB200_TAG="$v lone8 ticket x16" MGB200_LONE_WORLD=8 q
  MGB200_TAG="$v n1 ticket x16" q
done
for v in base pf; do
  export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  for h in 512 256 128; do
    MGB200_TAG="$v lone8 static x1 heavy>=$h" MGB200_HEAVY_MIN_DEGREE=$h MGB200_SELL_SCHED=static MGB200_SELL_WORK_ITEMS=$((148*32)) MGB200_LONE_WORLD=8 q
    MGB200_TAG="$v lone8 ticket x4 heavy>=$h" MGB200_HEAVY_MIN_DEGREE=$h MGB200_SELL_WORK_ITEMS=$((148*32*4)) MGB200_LONE_WORLD=8 q
  done
  MGB200_TAG="$v n1 heavy>=256" MGB200_HEAVY_MIN_DEGREE=256 q
done
