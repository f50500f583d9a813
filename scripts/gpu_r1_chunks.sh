#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/chunks.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_pagerank.py tests/test_gpu_multi.py tests/test_gpu_module.py -m gpu -x -q -k "not scale26 and not scale22" --tb=short 2>&1 | tail -4 | tee -a $O
MGB200_SELL_CHUNKS=3 timeout 900 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "not scale26 and not scale22" --tb=short 2>&1 | tail -2 | tee -a $O
runN() { N=$1; SC=$2; tag=$3; shift 3
  if [ "$N" = 1 ]; then env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $SC --steps 3 --warmup 3 2>>gpurun_out/chunks.err | tee -a $O
  else env MGB200_TAG="$tag" "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2965$N bench.py --quick --gpus $N --scale $SC --steps 3 --warmup 3 2>>gpurun_out/chunks.err | tee -a $O; fi; }
runN 1 26 "n1 chunks=1"
runN 1 26 "n1 chunks=3" MGB200_SELL_CHUNKS=3
for k in 1 2 3 4 6; do runN 2 26 "n2 s26 chunks=$k" MGB200_SELL_CHUNKS=$k; done
for k in 1 3 6; do runN 2 24 "n2 s24 chunks=$k" MGB200_SELL_CHUNKS=$k; done
