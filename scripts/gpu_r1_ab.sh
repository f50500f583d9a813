#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/ab.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_pagerank.py tests/test_gpu_multi.py -m gpu -x -q -k "not scale26 and not scale22" --tb=short 2>&1 | tail -25 | tee -a $O
run() { tag=$1; shift; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale 26 --steps 3 --warmup 3 2>>gpurun_out/ab.err | tee -a $O; }
for i in 1 2; do
run "old $i" MGB200_LIBRARY=memgraph_b200/_build/variants/old/libmgb200_pagerank.so
run "new $i"
done
