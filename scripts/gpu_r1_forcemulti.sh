#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/forcemulti.txt; : > $O
timeout 300 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "rmat_small or config1 or golden" 2>&1 | tail -2 | tee -a $O
MGB200_FORCE_MULTI_PATH=1 timeout 300 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "rmat_small or config1" 2>&1 | tail -2 | tee -a $O
MGB200_TAG="n1 single path" timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
MGB200_TAG="n1 forced multi path" MGB200_FORCE_MULTI_PATH=1 timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
