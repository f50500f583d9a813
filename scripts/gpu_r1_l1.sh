#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/l1.txt; : > $O
run() { tag=$1; sc=$2; shift 2; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $sc --steps 3 --warmup 3 2>>gpurun_out/l1.err | tee -a $O; }
timeout 600 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "not scale26" 2>&1 | tail -2 | tee -a $O
for k in 0 8 16 24 28 48 96 256; do run "l1hot=${k}K" 26 MGB200_L1_HOT_K=$k; done
