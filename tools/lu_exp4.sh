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
run "getrf 16384 default (pipe, pct 20)" -- $PROD 16384 8
run "getrf 16384 pipe off" ATLB200_LU_PIPE=0 -- $PROD 16384 5
run "getrf 16384 pipe off pct100" ATLB200_LU_PIPE=0 ATLB200_LU_RESERVE2_PCT=100 -- $PROD 16384 5
run "getrf 16384 pipe pct 0" ATLB200_LU_RESERVE2_PCT=0 -- $PROD 16384 5
run "getrf 16384 pipe pct 10" ATLB200_LU_RESERVE2_PCT=10 -- $PROD 16384 5
run "getrf 16384 pipe pct 35" ATLB200_LU_RESERVE2_PCT=35 -- $PROD 16384 5
run "getrf 16384 pipe pct 50" ATLB200_LU_RESERVE2_PCT=50 -- $PROD 16384 5
run "getrf 16384 fused256" ATLB200_TRSM_FUSED_MAXN=256 -- $PROD 16384 5
run "getrf 16384 fused off" ATLB200_TRSM_FUSED=0 -- $PROD 16384 5
run "getrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
run "getrf 8192" -- $PROD 8192 20
run "getrf 8192 pipe off" ATLB200_LU_PIPE=0 -- $PROD 8192 5
run "getrf 4096" -- $PROD 4096 20
run "getrf 32768" -- $PROD 32768 3
