#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/chunks8.txt; : > $O
runN() { N=$1; SC=$2; tag=$3; shift 3
  env MGB200_TAG="$tag" "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2966$N bench.py --quick --gpus $N --scale $SC --steps 3 --warmup 3 2>>gpurun_out/chunks8.err | tee -a $O; }
runN 8 26 "n8 chunks=1" MGB200_SELL_CHUNKS=1
runN 8 26 "n8 chunks=3" MGB200_SELL_CHUNKS=3
runN 8 26 "n8 chunks=2" MGB200_SELL_CHUNKS=2
runN 4 26 "n4 chunks=2" MGB200_SELL_CHUNKS=2
