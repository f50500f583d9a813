#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/multi4c.txt; : > $O
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pagerank.py -m gpu -x -q -k "not scale26 and not scale22" 2>&1 | tail -3 | tee -a $O
runN() { N=$1; SC=$2; tag=$3; shift 3
  if [ "$N" = 1 ]; then env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $SC --steps 3 --warmup 3 2>>gpurun_out/multi4c.err | tee -a $O
  else env MGB200_TAG="$tag" "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --quick --gpus $N --scale $SC --steps 3 --warmup 3 2>>gpurun_out/multi4c.err | tee -a $O; fi; }
runN 1 26 "n1 s26"
runN 4 26 "n4 s26"
runN 4 26 "n4 s26 l1hot=0" MGB200_L1_HOT_K=0
runN 4 26 "n4 s26 l2hot=0" MGB200_L2_HOT_MB=0
runN 4 25 "n4 s25"
runN 2 26 "n2 s26"
