#!/bin/bash
# r2 call 18 (2 GPUs): streamed / chunked build on partitions, personalised PageRank, quick N=2 line with the final defaults
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c18; O=gpurun_out/c18/out.txt; : > $O
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_personalized.py -q -x -k "not (4- or 8-)" 2>&1 | tail -5 | tee -a $O
timeout 200 python -m pytest tests/test_gpu_pagerank.py -q -x -k "streamed or variants or abort or nan" 2>&1 | tail -3 | tee -a $O
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 2>>gpurun_out/c18/err.txt | tee gpurun_out/c18/bench_n2.json | cut -c1-300 | tee -a $O
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --scale 27 --quick --steps 2 --warmup 3 2>>gpurun_out/c18/err.txt | tee -a $O
tail -4 gpurun_out/c18/err.txt | cut -c1-250 | tee -a $O
