#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c11; O=gpurun_out/c11/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c11/err.txt | tee -a $O; }
for lw in 2 4; do
  MGB200_TAG="lone$lw ticket x16" MGB200_LONE_WORLD=$lw q
  MGB200_TAG="lone$lw ticket x4" MGB200_SELL_WORK_ITEMS=$((148*32*4)) MGB200_LONE_WORLD=$lw q
  MGB200_TAG="lone$lw static x1" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32)) MGB200_LONE_WORLD=$lw q
  MGB200_TAG="lone$lw static x2" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32*2)) MGB200_LONE_WORLD=$lw q
done
MGB200_TAG="n1 static x2" MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32*2)) q
MGB200_TAG="n1 ticket x32" MGB200_SELL_WORK_ITEMS=$((148*32*32)) q
MGB200_TAG="n1 ticket x64" MGB200_SELL_WORK_ITEMS=$((148*32*64)) q
