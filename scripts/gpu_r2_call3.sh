#!/bin/bash
# r2 call 3: is the N=1 slowdown (5.98 vs 4.0 ms) code or box?  what does the epilogue cost live, un-overlapped?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c3; O=gpurun_out/c3/out.txt; : > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader | tee -a $O
q() { timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c3/err.txt | tee -a $O; }
(cd scripts/_bin/r01tree && MGB200_TAG="r01 tree n1" timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>../../../gpurun_out/c3/err.txt) | tee -a $O
MGB200_TAG="head n1" q
MGB200_TAG="head n1 no-overlap" MGB200_OVERLAP_EPILOGUE=0 q
MGB200_TAG="head lone8 no-overlap" MGB200_OVERLAP_EPILOGUE=0 MGB200_LONE_WORLD=8 q
MGB200_TAG="head lone8" MGB200_LONE_WORLD=8 q
(cd scripts/_bin/r01tree && MGB200_TAG="r01 tree n1 again" timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>../../../gpurun_out/c3/err.txt) | tee -a $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader | tee -a $O
