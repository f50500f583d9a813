#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_multi.log
echo "== bench N=1 (scale ${SCALE:-24})"
timeout 600 python bench.py --scale ${SCALE:-24} --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.log; tail -2 $O/bench_n1.log; cat $O/bench_n1.json
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --scale ${SCALE:-24} --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.log; echo "rc=$?"; tail -5 $O/bench_n2.log; cat $O/bench_n2.json
