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
  env "${envs[@]}" timeout 600 $BIN "$@" 2>&1 | tail -n 70
}
run "getrf 16384" -- $PROD 16384 12
run "getrf 16384 pct 30" ATLB200_LU_RESERVE2_PCT=30 -- $PROD 16384 5
run "getrf 16384 pct 15" ATLB200_LU_RESERVE2_PCT=15 -- $PROD 16384 5
run "getrf 8192" -- $PROD 8192 20
run "getrf 4096" -- $PROD 4096 20
run "getrf 2048" -- $PROD 2048 10
run "getrf 32768" -- $PROD 32768 3
run "potrf 16384" -- $PROD 16384 5 potrf
run "getrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
