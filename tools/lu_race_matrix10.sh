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
run "getrf 16384 default" -- $PROD 16384 6
for pw in 64 128 256; do run "getrf 16384 pw_tall=$pw" ATLB200_LU_PW_TALL=$pw -- $PROD 16384 6; done
run "getrf 32768 default" -- $PROD 32768 3
run "getrf 32768 pw_tall=128" ATLB200_LU_PW_TALL=128 -- $PROD 32768 3
run "potrf 16384 default" -- $PROD 16384 6 potrf
for nb in 256 384; do run "potrf 16384 llt_nb=$nb" ATLB200_LLT_NB=$nb -- $PROD 16384 6 potrf; run "potrf 8192 llt_nb=$nb" ATLB200_LLT_NB=$nb -- $PROD 8192 6 potrf; done
run "potrf 8192 default" -- $PROD 8192 6 potrf
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25
echo "=== profiles"
bash tools/prof_r02.sh 2>&1 | tail -12
echo DONE
