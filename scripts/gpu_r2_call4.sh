#!/bin/bash
# r2 call 4: which part of the merged row epilogue costs 7x?  (variants: memgraph_b200/_build/variants/epi*)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c4; O=gpurun_out/c4/out.txt; : > $O
for v in epi0 epi1 epi2 epi3 epi4 epi5; do
  L=$PWD/memgraph_b200/_build/variants/$v/libmgb200_pagerank.so
  MGB200_LIBRARY=$L MGB200_TAG="$v n1 no-overlap" MGB200_OVERLAP_EPILOGUE=0 timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c4/err.txt | tee -a $O
  MGB200_LIBRARY=$L MGB200_TAG="$v lone8 no-overlap" MGB200_OVERLAP_EPILOGUE=0 MGB200_LONE_WORLD=8 timeout 200 python bench.py --quick --steps 2 --warmup 3 2>>gpurun_out/c4/err.txt | tee -a $O
done
L=$PWD/memgraph_b200/_build/variants/epi4/libmgb200_pagerank.so
MGB200_LIBRARY=$L MGB200_TAG="epi4 n1 overlap" timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c4/err.txt | tee -a $O
MGB200_LIBRARY=$L MGB200_TAG="epi4 lone8 overlap" MGB200_LONE_WORLD=8 timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c4/err.txt | tee -a $O
MGB200_LIBRARY=$L timeout 300 python -m pytest tests/test_gpu_pagerank.py -q -x 2>&1 | tail -3 | tee -a $O
