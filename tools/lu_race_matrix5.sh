#!/bin/bash
cd "$(dirname "$0")/.."
BIN=tools/bin/lu_race
PROD=math_atlas_b200/libatlas_b200.so
FORCED="ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_RESERVE=0 ATLB200_GEMM_CTA_CAP=1000000 ATLB200_LU_PW=0"
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $label :: ${envs[*]} :: $*"
  env "${envs[@]}" timeout 900 $BIN "$@" 2>&1 | tail -n 60
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for rep in 1 2; do
for v in prod rel2 rel3 oldrelease; do
  L=tools/bin/libatlas_b200_$v.so; [ $v = prod ] && L=$PROD
  run "gemm $v 16384" -- $L 16384 6 gemm
  run "gemm $v 4096" -- $L 4096 10 gemm
done
done
for v in rel2 rel3; do
  run "forced race $v 8192" $FORCED -- tools/bin/libatlas_b200_$v.so 8192 60
  run "forced race $v 16384" $FORCED -- tools/bin/libatlas_b200_$v.so 16384 25
done
for nb in 512 384 256; do
  run "getrf 16384 nb=$nb" ATLB200_LU_NB=$nb -- $PROD 16384 5
  run "getrf 8192 nb=$nb" ATLB200_LU_NB=$nb -- $PROD 8192 5
done
for pct in 0 15 25 40; do
  run "getrf 16384 reserve2 from $pct%" ATLB200_LU_RESERVE2_PCT=$pct -- $PROD 16384 5
  run "getrf 8192 reserve2 from $pct%" ATLB200_LU_RESERVE2_PCT=$pct -- $PROD 8192 5
done
run "getrf 16384 nb=256 reserve2 25%" ATLB200_LU_NB=256 ATLB200_LU_RESERVE2_PCT=25 -- $PROD 16384 5
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -40
echo DONE
