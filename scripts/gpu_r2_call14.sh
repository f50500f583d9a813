#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c14; O=gpurun_out/c14/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c14/err.txt | tee -a $O; }
for c in 0 1 2 4 8; do
  MGB200_TAG="lone8 static x1 cost=$c" MGB200_SLICE_COST=$c MGB200_LONE_WORLD=8 q
  MGB200_TAG="lone8 static x2 cost=$c" MGB200_SLICE_COST=$c MGB200_SELL_WORK_ITEMS=$((148*64)) MGB200_LONE_WORLD=8 q
  MGB200_TAG="lone8 ticket x4 cost=$c" MGB200_SLICE_COST=$c MGB200_SELL_MODE=0 MGB200_SELL_WORK_ITEMS=$((148*32*4)) MGB200_LONE_WORLD=8 q
  MGB200_TAG="lone4 static x2 cost=$c" MGB200_SLICE_COST=$c MGB200_LONE_WORLD=4 q
  MGB200_TAG="n1 ticket x16 cost=$c" MGB200_SLICE_COST=$c q
done
