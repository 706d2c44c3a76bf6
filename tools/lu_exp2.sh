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
run "getrf 16384 new" -- $PROD 16384 6
run "getrf 16384 old laswp" ATLB200_LASWP_PAR=0 -- $PROD 16384 4
run "getrf 16384 old trsm" ATLB200_TRSM_FUSED=0 -- $PROD 16384 4
run "getrf 16384 both old" ATLB200_TRSM_FUSED=0 ATLB200_LASWP_PAR=0 -- $PROD 16384 4
run "getrf 16384 reserve2 from 30%" ATLB200_LU_RESERVE2_PCT=30 -- $PROD 16384 4
run "getrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
run "getrf 8192" -- $PROD 8192 6
run "getrf 4096" -- $PROD 4096 6
run "getrf 2048" -- $PROD 2048 6
run "getrf 32768" -- $PROD 32768 2
