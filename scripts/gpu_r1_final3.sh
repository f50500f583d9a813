#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/final3; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
MGB200_TAG="default" timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee $O/quick.txt
MGB200_TAG="plainL1 hot=0" MGB200_L1_HOT_K=-1 MGB200_L2_HOT_MB=0 timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O/quick.txt
