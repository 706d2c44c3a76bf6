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
run "getrf 16384 fused<=128" -- $PROD 16384 5
run "getrf 16384 fused<=256" ATLB200_TRSM_FUSED_MAXN=256 -- $PROD 16384 5
run "getrf 16384 fused off" ATLB200_TRSM_FUSED=0 -- $PROD 16384 5
run "getrf 8192 fused<=128" -- $PROD 8192 5
run "getrf 8192 fused off" ATLB200_TRSM_FUSED=0 -- $PROD 8192 5
run "getrf 4096 fused<=128" -- $PROD 4096 5
run "getrf 4096 fused off" ATLB200_TRSM_FUSED=0 -- $PROD 4096 5
run "getrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
