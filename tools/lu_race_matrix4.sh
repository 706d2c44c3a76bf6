#!/bin/bash
cd "$(dirname "$0")/.."
BIN=tools/bin/lu_race
PROD=math_atlas_b200/libatlas_b200.so
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $label :: ${envs[*]} :: $*"
  env "${envs[@]}" timeout 900 $BIN "$@" 2>&1 | tail -n 120
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for rep in 1 2; do
for v in prod oldrelease; do
  L=tools/bin/libatlas_b200_$v.so; [ $v = prod ] && L=$PROD
  run "gemm $v 16384" -- $L 16384 6 gemm
  run "gemm $v 8192" -- $L 8192 10 gemm
done
done
run "phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
run "phase trace 8192" ATLB200_LU_TRACE=2 -- $PROD 8192 2
run "phase trace 16384 reserve0" ATLB200_LU_TRACE=2 ATLB200_LU_RESERVE=0 -- $PROD 16384 2
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -40
echo DONE
