#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c9; O=gpurun_out/c9/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c9/err.txt | tee -a $O; }
for v in t_u8b4pf t_u4b6pf; do
  export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  for mode in 0 1 2; do
    MGB200_TAG="$v lone8 mode$mode" MGB200_SELL_MODE=$mode MGB200_LONE_WORLD=8 q
    MGB200_TAG="$v n1 mode$mode" MGB200_SELL_MODE=$mode q
  done
  MGB200_TAG="$v lone8 mode1 x1" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32)) MGB200_LONE_WORLD=8 q
  MGB200_TAG="$v lone8 mode1 x1.5" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*48)) MGB200_LONE_WORLD=8 q
done
export MGB200_LIBRARY=$PWD/memgraph_b200/_build/variants/t_u8b4pf/libmgb200_pagerank.so
MGB200_SELL_MODE=0 MGB200_LONE_WORLD=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sell_rows' -s 6 -c 1 -o gpurun_out/c9/lone8_m0 -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c9/ncu0.log 2>&1
MGB200_SELL_MODE=2 MGB200_LONE_WORLD=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sell_rows' -s 6 -c 1 -o gpurun_out/c9/lone8_m2 -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c9/ncu2.log 2>&1
