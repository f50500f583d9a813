#!/bin/bash
# Round-2 first run: validate everything that was written without a GPU at the end of round 1
# (global-order labelling, push need-mask, Katz).  1 GPU: full -m gpu suite + quick bench;  N GPUs (gpurun --gpus N):
# the partitioned tests over both labellings x push mask, then a quick A/B of the four combinations.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/r2_validate.txt; : > $O
NG=$(python -c "import torch; print(torch.cuda.device_count())" 2>/dev/null || echo 1)
echo "gpus $NG" | tee -a $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a $O
MGB200_TAG="n1 default" timeout 200 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
# what ONE GPU of an 8-GPU run does, without the other seven (profiling mode; add ncu in front of python when needed):
for v in "" "MGB200_IDX_FLAGS=0" "MGB200_LABELLING=global"; do
  env $v MGB200_TAG="lone 1/8 partition $v" MGB200_LONE_WORLD=8 timeout 200 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
done
if [ "$NG" -ge 2 ]; then
  for lab in dealt global; do for pm in 0 1; do
    MGB200_TAG="n$NG labelling=$lab push_mask=$pm" MGB200_LABELLING=$lab MGB200_PUSH_MASK=$pm timeout 300 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $NG --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
  done; done
  MGB200_TAG="n$NG push=copy" MGB200_PUSH=copy timeout 300 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O
fi
