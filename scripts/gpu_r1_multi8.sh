#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/topo8.txt 2>&1
echo "== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_multi8.log
for N in 8 4 2; do
echo "== bench N=$N scale 26"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > $O/bench26_n$N.json 2> $O/bench26_n$N.log; echo "rc=$?"; grep -E "rank [0-9]\]|rror" $O/bench26_n$N.log | head -10; cat $O/bench26_n$N.json
done
echo "== bench N=1 scale 26"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench26_n1.json 2> $O/bench26_n1.log; cat $O/bench26_n1.json
