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
  env "${envs[@]}" timeout 900 $BIN "$@" 2>&1 | tail -n 60
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
run "getrf determinism 16384" -- $PROD 16384 30
run "getrf determinism 8192" -- $PROD 8192 40
run "getrf determinism 4096" -- $PROD 4096 40
run "getrf determinism 2048" -- $PROD 2048 20
run "getrf forced 8192" ATLB200_LU_RESERVE=0 ATLB200_GEMM_CTA_CAP=1000000 -- $PROD 8192 40
run "getrf 32768" -- $PROD 32768 4
run "getrf 24576" -- $PROD 24576 4
run "potrf determinism 16384" -- $PROD 16384 30 potrf
run "potrf determinism 8192" -- $PROD 8192 40 potrf
for r in 0 4 16; do run "potrf 16384 llt_reserve=$r" ATLB200_LLT_RESERVE=$r -- $PROD 16384 6 potrf; done
run "potrf 32768" -- $PROD 32768 3 potrf
run "getrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
run "potrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2 potrf
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25
echo DONE
