#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c12; O=gpurun_out/c12/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c12/err.txt | tee -a $O; }
timeout 300 python -m pytest tests/test_gpu_pagerank.py -q -x 2>&1 | tail -2 | tee -a $O
for k in 1 0; do
  MGB200_TAG="n1 keep=$k" MGB200_KEEP_HOT_STORES=$k q
  MGB200_TAG="n1 keep=$k no-overlap" MGB200_OVERLAP_EPILOGUE=0 MGB200_KEEP_HOT_STORES=$k q
  MGB200_TAG="lone8 static x1 keep=$k" MGB200_KEEP_HOT_STORES=$k MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32)) MGB200_LONE_WORLD=8 q
  MGB200_TAG="lone4 static x2 keep=$k" MGB200_KEEP_HOT_STORES=$k MGB200_SELL_MODE=1 MGB200_SELL_WORK_ITEMS=$((148*32*2)) MGB200_LONE_WORLD=4 q
done
for mb in 32 48 96; do MGB200_TAG="n1 keep=1 hot=${mb}MB" MGB200_L2_HOT_MB=$mb q; done
