#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/sweep2.txt; : > $O
V=memgraph_b200/_build/variants
run() { tag=$1; shift; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 2>/dev/null | tee -a $O; }
for v in base noidxhint noepihint nogather allplain gatherlast gatherlast_noepi b3 b2; do run "variant=$v" MGB200_LIBRARY=$V/$v/libmgb200_pagerank.so; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_rows -s 3 -c 1 -o gpurun_out/prof_sell_s26_v2 -f \
  python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_v2.log 2>&1; echo "ncu rc=$?"
