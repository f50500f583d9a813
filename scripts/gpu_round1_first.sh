#!/bin/bash
# First GPU session: smoke, parity tests, bench at two scales, launch list + one ncu --set full capture.
# Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvidia-smi.txt 2>&1
lscpu > $O/lscpu.txt 2>&1; free -g > $O/free.txt 2>&1
echo "== smoke" | tee $O/smoke.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
tail -5 $O/smoke.log
echo "== pytest gpu (small first)"
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
echo "== bench scale 22"
timeout 600 python bench.py --scale 22 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_s22.json 2> $O/bench_s22.log; echo "rc=$?"
cat $O/bench_s22.json; tail -3 $O/bench_s22.log
echo "== bench scale 26"
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_s26.json 2> $O/bench_s26.log; echo "rc=$?"
cat $O/bench_s26.json; tail -3 $O/bench_s26.log
echo "== ncu launch list (scale 24)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_s24.csv \
  python bench.py --scale 24 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full on sell kernel (scale 26)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sell_rows -s 3 -c 2 -o $O/prof_sell_s26 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_bench.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heavy_segments -s 3 -c 1 -o $O/prof_heavy_s26 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_heavy.log 2>&1; echo "rc=$?"
ls -la $O
