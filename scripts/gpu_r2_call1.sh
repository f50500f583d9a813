#!/bin/bash
# r2 call 1 (1 GPU): validate the merged tree (Katz, labelling, mask), N=1 quick bench, lone 1/8 partition timings + ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c1; O=gpurun_out/c1/out.txt; : > $O
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | tee -a $O
nproc | tee -a $O; free -g | head -2 | tee -a $O
timeout 900 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -45 | tee -a $O
MGB200_TAG="n1 default" timeout 200 python bench.py --quick --steps 3 --warmup 3 2>gpurun_out/c1/n1.err | tee -a $O
for v in "MGB200_X=0" "MGB200_IDX_FLAGS=0" "MGB200_LABELLING=global"; do
  env $v MGB200_TAG="lone 1/8 $v" MGB200_LONE_WORLD=8 timeout 200 python bench.py --quick --steps 3 --warmup 3 2>gpurun_out/c1/lone.err | tee -a $O
done
for v in "MGB200_X=0" "MGB200_LABELLING=global"; do
  env $v MGB200_TAG="lone 1/4 $v" MGB200_LONE_WORLD=4 timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c1/lone.err | tee -a $O
done
# ncu: one 1/8 partition, the two gather kernels (dealt labelling, default flags)
MGB200_LONE_WORLD=8 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sell_rows|heavy_segments|sell_epilogue' -s 12 -c 3 \
  -o gpurun_out/c1/lone8 -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c1/ncu.log 2>&1
tail -3 gpurun_out/c1/ncu.log | tee -a $O
