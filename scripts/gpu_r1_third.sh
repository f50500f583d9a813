#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
timeout 600 scripts/_bin/l2_bench > $O/l2_bench.txt 2>&1; echo "rc=$?"
cat $O/l2_bench.txt
timeout 600 python -m pytest tests/test_gpu_module.py -m gpu -x -q > $O/pytest_gpu_module.log 2>&1; echo "rc=$?"
tail -5 $O/pytest_gpu_module.log
