#!/bin/bash
# Round-1 final single-GPU pass: smoke, full parity suite, sanitizer on a small case, default bench (both arms),
# BFS bench, launch list + one --set full capture of the dominant kernel.  Outputs -> gpurun_out/final1/
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/final1; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $O/cpu.txt; free -g >> $O/cpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.txt
echo "== compute-sanitizer memcheck (RMAT scale-12 PageRank + BFS)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import numpy as np, memgraph_b200 as mg
from memgraph_b200 import bfs
f,t = mg.rmat_edges_host(12, 16<<12)
g = mg.PageRankGraph.from_arrays(1<<12, f, t); r,s = g.run(20,0.85,0.0); print('pagerank sum', r.sum(), s.iterations); g.close()
b = bfs.BfsGraph(1<<12, f, t); d,st = b.distances(0); print('bfs reached', st['reached']); b.close()
" > $O/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.txt; tail -4 $O/sanitizer_memcheck.txt
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.log; cat $O/bench_reference.json
echo "== bench default"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.log; cat $O/bench_default.json; tail -2 $O/bench_default.log
echo "== bench bfs"; timeout 600 python bench.py --workload bfs --steps 5 > $O/bench_bfs.json 2> $O/bench_bfs.log; cat $O/bench_bfs.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_s26.csv \
  python bench.py --quick --steps 1 --warmup 3 > $O/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full: sell_rows + heavy_segments + sell_epilogue"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"sell_rows|heavy_segments|sell_epilogue" -s 9 -c 3 -o $O/prof_final_s26 -f \
  python bench.py --quick --steps 1 --warmup 3 > $O/ncu_full.log 2>&1; echo "rc=$?"
ls -la $O
