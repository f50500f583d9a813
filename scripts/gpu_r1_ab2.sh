#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/ab2.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_pagerank.py tests/test_gpu_multi.py -m gpu -x -q -k "not scale26 and not scale22" --tb=short 2>&1 | tail -4 | tee -a $O
runN() { N=$1; SC=$2; tag=$3; shift 3
  if [ "$N" = 1 ]; then env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $SC --steps 3 --warmup 3 2>>gpurun_out/ab2.err | tee -a $O
  else env MGB200_TAG="$tag" "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N bench.py --quick --gpus $N --scale $SC --steps 3 --warmup 3 2>>gpurun_out/ab2.err | tee -a $O; fi; }
runN 1 26 "n1 old" MGB200_LIBRARY=memgraph_b200/_build/variants/old/libmgb200_pagerank.so
runN 1 26 "n1 new"
runN 1 26 "n1 new plainL1" MGB200_L1_HOT_K=-1
if [ "${NG:-1}" -ge 2 ]; then
runN $NG 26 "n$NG new (aware, l1=16K)"
runN $NG 26 "n$NG aware plainL1" MGB200_L1_HOT_K=-1
runN $NG 26 "n$NG legacy plainL1" MGB200_L1_HOT_K=-1 MGB200_MULTI_AWARE=0
runN $NG 26 "n$NG legacy l1=16K" MGB200_MULTI_AWARE=0
runN $NG 26 "n$NG aware plainL1 l2hot=32" MGB200_L1_HOT_K=-1 MGB200_L2_HOT_MB=32
fi
