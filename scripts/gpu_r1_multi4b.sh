#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; O=gpurun_out/multi4b.txt; : > $O
runN() { N=$1; SC=$2; tag=$3; shift 3
  if [ "$N" = 1 ]; then env MGB200_TAG="$tag" "$@" timeout 300 python bench.py --quick --scale $SC --steps 3 --warmup 3 2>>gpurun_out/multi4b.err | tee -a $O
  else env MGB200_TAG="$tag" "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --quick --gpus $N --scale $SC --steps 3 --warmup 3 2>>gpurun_out/multi4b.err | tee -a $O; fi; }
runN 4 26 "n4 s26"
runN 4 25 "n4 s25"
runN 4 25 "n4 s25 no-overlap" MGB200_OVERLAP_EPILOGUE=0
runN 4 24 "n4 s24"
runN 2 25 "n2 s25"
