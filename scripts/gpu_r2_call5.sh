#!/bin/bash
# r2 call 5: new row epilogue + ticketed SELL kernel: parity, sanitizer, N=1 and lone-partition timings
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c5; O=gpurun_out/c5/out.txt; : > $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee -a $O
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee -a $O
timeout 300 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $O
q() { timeout 200 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/c5/err.txt | tee -a $O; }
MGB200_TAG="n1" q
MGB200_TAG="n1 no-overlap" MGB200_OVERLAP_EPILOGUE=0 q
MGB200_TAG="lone8" MGB200_LONE_WORLD=8 q
MGB200_TAG="lone8 no-overlap" MGB200_OVERLAP_EPILOGUE=0 MGB200_LONE_WORLD=8 q
MGB200_TAG="lone8 global" MGB200_LABELLING=global MGB200_LONE_WORLD=8 q
MGB200_TAG="lone4" MGB200_LONE_WORLD=4 q
MGB200_TAG="lone2" MGB200_LONE_WORLD=2 q
for w in 4 8 32; do MGB200_TAG="n1 work x$w" MGB200_SELL_WORK_ITEMS=$((148*32*w)) q; MGB200_TAG="lone8 work x$w" MGB200_SELL_WORK_ITEMS=$((148*32*w)) MGB200_LONE_WORLD=8 q; done
