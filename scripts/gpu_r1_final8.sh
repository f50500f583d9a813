#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/final8; mkdir -p $O
echo "== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_multi.txt
for N in 8 4 2; do
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.log; echo "rc=$?"; cat $O/bench_n$N.json
done
echo "== bench N=1"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.log; cat $O/bench_n1.json
