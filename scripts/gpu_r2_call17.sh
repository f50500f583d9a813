#!/bin/bash
# r2 call 17 (1 GPU): bank the records -- full GPU suite, contract line with e2e_call, katz / bfs lines, reference arm,
# ncu launch list + full captures (plain vs hot-table SELL kernel, heavy, epilogue)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/c17; O=gpurun_out/c17/out.txt; : > $O
timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -16 | tee -a $O
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c17/bench_default.json 2> gpurun_out/c17/bench_default.err; tail -c 600 gpurun_out/c17/bench_default.json | tee -a $O
timeout 300 python bench.py --workload katz --steps 3 --warmup 3 > gpurun_out/c17/bench_katz.json 2>> gpurun_out/c17/err.txt; cut -c1-400 gpurun_out/c17/bench_katz.json | tee -a $O
timeout 300 python bench.py --workload bfs --steps 5 --warmup 3 > gpurun_out/c17/bench_bfs.json 2>> gpurun_out/c17/err.txt; cut -c1-400 gpurun_out/c17/bench_bfs.json | tee -a $O
MGB200_REF_BUDGET_S=70 timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c17/bench_reference.json 2>> gpurun_out/c17/err.txt; cut -c1-300 gpurun_out/c17/bench_reference.json | tee -a $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c17/launches_s26.csv python bench.py --quick --steps 1 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sell_rows|heavy_segments|row_epilogue' -s 15 -c 4 -o gpurun_out/c17/n1_full -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c17/ncu1.log 2>&1
MGB200_SMEM_TABLE_KB=128 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sell_rows' -s 4 -c 1 -o gpurun_out/c17/n1_table128 -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c17/ncu2.log 2>&1
MGB200_LONE_WORLD=8 timeout 400 ncu --set full --clock-control none -k regex:'sell_rows|heavy_segments|row_epilogue' -s 15 -c 4 -o gpurun_out/c17/lone8_full -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/c17/ncu3.log 2>&1
ls -la gpurun_out/c17 | tee -a $O
