#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c24; O=gpurun_out/c24/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c24/err.txt | tee -a $O; }
MGB200_SELL_POOL_PERMILLE=200 timeout 200 python -m pytest tests/test_gpu_pagerank.py -q -x -k "config2 or streamed" 2>&1 | tail -2 | tee -a $O
for lw in 8 4 2; do
  MGB200_TAG="lone$lw base" MGB200_LONE_WORLD=$lw q
  for pm in 100 200 350; do for pi in 1 2 4; do
    MGB200_TAG="lone$lw pool=$pm items=$pi" MGB200_SELL_MODE=1 MGB200_SELL_POOL_PERMILLE=$pm MGB200_SELL_POOL_ITEMS=$pi MGB200_LONE_WORLD=$lw q
  done; done
done
MGB200_TAG="n1 base" q
MGB200_TAG="n1 static pool=200 items=4" MGB200_SELL_MODE=1 MGB200_SELL_POOL_PERMILLE=200 MGB200_SELL_POOL_ITEMS=4 q
MGB200_TAG="n1 static pool=350 items=8" MGB200_SELL_MODE=1 MGB200_SELL_POOL_PERMILLE=350 MGB200_SELL_POOL_ITEMS=8 q
