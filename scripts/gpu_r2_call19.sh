#!/bin/bash
# r2 call 19 (8 GPUs): final contract lines at N=8 / N=4 (with the parity field), and RMAT scale-28 (4.3 B edges) on 8 GPUs
# built from the generator stream
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c19; O=gpurun_out/c19/out.txt; : > $O
tr() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $1 "${@:2}" 2>>gpurun_out/c19/err.txt; }
tr 8 --steps 5 --warmup 3 | tee gpurun_out/c19/bench_n8.json | cut -c1-200 | tee -a $O
tr 4 --steps 5 --warmup 3 | tee gpurun_out/c19/bench_n4.json | cut -c1-200 | tee -a $O
tr 8 --scale 28 --quick --steps 2 --warmup 3 | tee gpurun_out/c19/quick_s28_n8.json | tee -a $O
grep "RMAT scale-28" gpurun_out/c19/err.txt | cut -c1-300 | tee -a $O
