#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c6; O=gpurun_out/c6/out.txt; : > $O
q() { timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c6/err.txt | tee -a $O; }
for w in 1 2 4 16; do for sched in static ticket; do
  MGB200_TAG="lone8 $sched x$w" MGB200_SELL_SCHED=$sched MGB200_SELL_WORK_ITEMS=$((148*32*w)) MGB200_LONE_WORLD=8 q
done; done
for w in 1 4 16; do MGB200_TAG="n1 static x$w" MGB200_SELL_SCHED=static MGB200_SELL_WORK_ITEMS=$((148*32*w)) q; done
MGB200_TAG="lone8 global static x4" MGB200_LABELLING=global MGB200_SELL_SCHED=static MGB200_SELL_WORK_ITEMS=$((148*32*4)) MGB200_LONE_WORLD=8 q
MGB200_LONE_WORLD=8 MGB200_SELL_WORK_ITEMS=$((148*32*4)) timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sell_rows' -s 6 -c 1 -o gpurun_out/c6/lone8_ticket -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c6/ncu.log 2>&1
