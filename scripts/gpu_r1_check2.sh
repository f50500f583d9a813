#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/check2.txt; : > $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $O
run() { tag=$1; sc=$2; shift 2; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $sc --steps 3 --warmup 3 2>>gpurun_out/check2.err | tee -a $O; }
run "n1 s26" 26
run "n1 s26 no-overlap" 26 MGB200_OVERLAP_EPILOGUE=0
run "n1 s23" 23
run "n1 s22" 22
