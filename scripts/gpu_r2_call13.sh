#!/bin/bash
# r2 call 13 (8 GPUs): multi-GPU parity tests over every surviving mode, A/B of labelling x push mask x copy push at N=8,
# then the contract line (with the parity field) at N=8/4/2.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c13; O=gpurun_out/c13/out.txt; : > $O
NG=$(nvidia-smi -L | wc -l); echo "gpus $NG" | tee -a $O
timeout 480 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -6 | tee -a $O
tr() { timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $1 "${@:2}" 2>>gpurun_out/c13/err.txt | tee -a $O; }
for cfg in "MGB200_X=0" "MGB200_PUSH_MASK=1" "MGB200_LABELLING=global" "MGB200_LABELLING=global MGB200_PUSH_MASK=1" "MGB200_PUSH=copy"; do
  ( export $cfg; export MGB200_TAG="n8 $cfg"; tr 8 --quick --steps 2 --warmup 3 )
done
tr 8 --steps 5 --warmup 3 
tr 4 --steps 5 --warmup 3
tr 2 --steps 5 --warmup 3
