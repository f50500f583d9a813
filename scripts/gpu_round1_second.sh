#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out
echo "== gather bench"
timeout 900 scripts/_bin/gather_bench > $O/gather_bench.txt 2>&1; echo "rc=$?"
cat $O/gather_bench.txt
echo "== pytest gpu module"
timeout 900 python -m pytest tests/test_gpu_module.py -m gpu -x -q > $O/pytest_gpu_module.log 2>&1; echo "rc=$?"
tail -15 $O/pytest_gpu_module.log
