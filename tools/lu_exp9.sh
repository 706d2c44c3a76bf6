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
for a in 0 1 2 3; do run "getrf 16384 aux_after=$a" ATLB200_LU_AUX_AFTER=$a -- $PROD 16384 6; done
for a in 0 2; do run "getrf 8192 aux_after=$a" ATLB200_LU_AUX_AFTER=$a -- $PROD 8192 8; done
for a in 0 2; do run "getrf 32768 aux_after=$a" ATLB200_LU_AUX_AFTER=$a -- $PROD 32768 2; done
