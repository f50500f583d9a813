#!/bin/bash
# A/B on one B200: hot flags baked into the stored indices (MGB200_IDX_FLAGS=1) against the range-policy single path
# and the per-gather owner lookup (MGB200_FORCE_MULTI_PATH=1, MGB200_IDX_FLAGS=0).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/idxflags.txt; : > $O
MGB200_IDX_FLAGS=1 timeout 120 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "rmat_small or config1 or golden" 2>&1 | tail -2 | tee -a $O
MGB200_TAG="n1 range policy (default)" timeout 100 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
MGB200_TAG="n1 index flags" MGB200_IDX_FLAGS=1 timeout 100 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
MGB200_TAG="n1 owner lookup" MGB200_IDX_FLAGS=0 MGB200_FORCE_MULTI_PATH=1 timeout 100 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
