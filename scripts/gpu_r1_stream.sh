#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/stream1.txt; : > $O
V=memgraph_b200/_build/variants
run() { tag=$1; shift; env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 2>>gpurun_out/stream1.err | tee -a $O; }
echo "== parity (stream kernel)" | tee -a $O
timeout 600 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "not scale26" 2>&1 | tail -5 | tee -a $O
echo "== parity (rows kernel)" | tee -a $O
MGB200_SELL_KERNEL=rows timeout 600 python -m pytest tests/test_gpu_pagerank.py -m gpu -x -q -k "not scale26 and not scale22" 2>&1 | tail -3 | tee -a $O
run "rows" MGB200_SELL_KERNEL=rows
run "stream default"
for v in s_w8_i3_v6 s_w8_i2_v8 s_w12_i2_v4 s_w16_i2_v3 s_w4_i4_v8 s_w8_i3_v2; do run "variant=$v" MGB200_LIBRARY=$V/$v/libmgb200_pagerank.so; done
for it in 1184 2368 9472 18944; do run "items=$it" MGB200_SELL_ITEMS=$it; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_stream -s 3 -c 1 -o gpurun_out/prof_stream_s26 -f \
  python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_stream.log 2>&1; echo "ncu rc=$?"
