#!/bin/bash
# r2 call 20 (1 GPU): everything written since the last GPU run -- edge weights, cuGraph stand-in modules, streamed build
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c20; O=gpurun_out/c20/out.txt; : > $O
timeout 400 python -m pytest tests/test_gpu_personalized.py tests/test_cugraph_modules.py tests/test_gpu_module.py tests/test_bfs_module.py tests/test_katz_module.py -q -m gpu 2>&1 | tail -8 | tee -a $O
timeout 200 python -m pytest tests/test_gpu_pagerank.py -q -x -k "streamed or variants or golden or edge_case" 2>&1 | tail -3 | tee -a $O
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_personalized.py -q -k "fixtures or weights" 2>&1 | tail -4 | tee -a $O
MGB200_TAG="n1" timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c20/err.txt | tee -a $O
