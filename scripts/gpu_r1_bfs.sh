#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/bfs.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_bfs.py -m gpu -x -q -s 2>&1 | tail -15 | tee -a $O
bash scripts/gpu_r1_l1.sh
