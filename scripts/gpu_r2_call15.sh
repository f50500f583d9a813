#!/bin/bash
# r2 call 15 (8 GPUs): the push kernel as a small high-priority grid next to the heavy-row kernel
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c15; O=gpurun_out/c15/out.txt; : > $O
tr() { timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $1 "${@:2}" 2>>gpurun_out/c15/err.txt | tee -a $O; }
for cfg in "MGB200_PUSH_CTAS=1" "MGB200_PUSH_CTAS=2" "MGB200_PUSH_CTAS=4" "MGB200_PUSH_CTAS=0" "MGB200_PUSH_CTAS=1 MGB200_PUSH_MASK=1" "MGB200_PUSH_CTAS=2 MGB200_PUSH_MASK=1"; do
  ( export $cfg; export MGB200_TAG="n8 $cfg"; tr 8 --quick --steps 2 --warmup 3 )
done
( export MGB200_PUSH_CTAS=1 MGB200_TAG="n4 ctas=1"; tr 4 --quick --steps 2 --warmup 3 )
( export MGB200_PUSH_CTAS=2 MGB200_TAG="n4 ctas=2"; tr 4 --quick --steps 2 --warmup 3 )
( export MGB200_PUSH_CTAS=0 MGB200_TAG="n4 ctas=0"; tr 4 --quick --steps 2 --warmup 3 )
( export MGB200_PUSH_CTAS=1 MGB200_TAG="n2 ctas=1"; tr 2 --quick --steps 2 --warmup 3 )
( export MGB200_PUSH_CTAS=0 MGB200_TAG="n2 ctas=0"; tr 2 --quick --steps 2 --warmup 3 )
timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c15/err.txt | tee -a $O
