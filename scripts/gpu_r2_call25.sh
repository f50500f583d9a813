#!/bin/bash
# r2 call 25 (8 GPUs, short): shift rows from the SELL class into the heavy class so that the second phase hides the push
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c25; O=gpurun_out/c25/out.txt; : > $O
tr() { timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --quick --steps 2 --warmup 3 2>>gpurun_out/c25/err.txt | tee -a $O; }
for h in 256 512 128; do ( export MGB200_HEAVY_MIN_DEGREE=$h MGB200_TAG="n8 heavy>=$h"; tr ); done
