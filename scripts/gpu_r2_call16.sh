#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c16; O=gpurun_out/c16/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c16/err.txt | tee -a $O; }
timeout 400 python -m pytest tests/test_gpu_pagerank.py tests/test_gpu_katz.py -q -x 2>&1 | tail -3 | tee -a $O
for kb in 0 64 96 128 160 192 208 224; do
  MGB200_TAG="n1 table=${kb}KB" MGB200_SMEM_TABLE_KB=$kb q
done
for kb in 128 192; do for l1 in 0 16 32 48; do
  MGB200_TAG="n1 table=${kb}KB l1hot=+${l1}K" MGB200_SMEM_TABLE_KB=$kb MGB200_L1_HOT_K=$((kb/8 + l1)) q
done; done
for kb in 0 96 160 208; do
  MGB200_TAG="lone8 table=${kb}KB" MGB200_SMEM_TABLE_KB=$kb MGB200_LONE_WORLD=8 q
done
