#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/check1.txt; : > $O
timeout 900 python -m pytest tests -m gpu -x -q -k "not scale26" 2>&1 | tail -3 | tee -a $O
run() { tag=$1; sc=$2; shift 2; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $sc --steps 3 --warmup 3 2>>gpurun_out/check1.err | tee -a $O; }
run "n1 s26" 26
run "n1 s23" 23
run "n1 s23 seg4096" 23 MGB200_SEGMENT_EDGES=4096
run "n1 s26 heavy_min=512" 26 MGB200_HEAVY_MIN_DEGREE=512
run "n1 s26 heavy_min=2048" 26 MGB200_HEAVY_MIN_DEGREE=2048
run "n1 s26 hot=48" 26 MGB200_L2_HOT_MB=48
run "n1 s26 hot=96" 26 MGB200_L2_HOT_MB=96
