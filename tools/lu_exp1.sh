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
run "getrf 16384" -- $PROD 16384 4
run "getrf 16384 reserve2 from start" ATLB200_LU_RESERVE2_PCT=0 -- $PROD 16384 4
run "getrf 16384 reserve2 from 30%" ATLB200_LU_RESERVE2_PCT=30 -- $PROD 16384 4
run "getrf 16384 pw_tall 256" ATLB200_LU_PW_TALL=256 -- $PROD 16384 4
run "getrf 16384 pw 128" ATLB200_LU_PW=128 -- $PROD 16384 4
run "getrf 16384 pw 256" ATLB200_LU_PW=256 -- $PROD 16384 4
run "getrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
run "getrf 8192" -- $PROD 8192 4
run "getrf 4096" -- $PROD 4096 4
run "getrf 2048" -- $PROD 2048 4
run "potrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2 potrf
