#!/bin/bash
# r2 call 22 (1 GPU): final validation of the tree as committed -- full GPU suite, smoke, contract line, ncu captures
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c22; O=gpurun_out/c22/out.txt; : > $O
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 | tee -a $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c22/bench_default.json 2> gpurun_out/c22/bench_default.err; cut -c1-250 gpurun_out/c22/bench_default.json | tee -a $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c22/launches_s26.csv python bench.py --quick --steps 1 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sell_rows|heavy_segments|row_epilogue' -s 15 -c 4 -o gpurun_out/c22/n1_full -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c22/ncu1.log 2>&1
MGB200_LONE_WORLD=8 timeout 400 ncu --set full --clock-control none -k regex:'sell_rows|heavy_segments|row_epilogue' -s 15 -c 4 -o gpurun_out/c22/lone8_full -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c22/ncu3.log 2>&1
ls -la gpurun_out/c22 | tee -a $O
