#!/bin/bash
cd "$(dirname "$0")/.."
BIN=tools/bin/lu_race
PROD=math_atlas_b200/libatlas_b200.so
DIAG=tools/bin/libatlas_b200_diag.so
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
run "sched3 determinism 16384" -- $PROD 16384 30
run "sched3 determinism 8192" -- $PROD 8192 40
run "sched3 determinism 4096" -- $PROD 4096 40
run "sched3 forced 8192" ATLB200_LU_RESERVE=0 ATLB200_GEMM_CTA_CAP=1000000 -- $PROD 8192 40
run "sched2 16384" ATLB200_LU_SCHED=2 -- $PROD 16384 5
run "sched2 8192" ATLB200_LU_SCHED=2 -- $PROD 8192 5
for pw in 64 128 256; do
  run "sched3 16384 pw=$pw" ATLB200_LU_PW=$pw -- $PROD 16384 5
  run "sched3 8192 pw=$pw" ATLB200_LU_PW=$pw -- $PROD 8192 5
done
for pct in 10 20 30; do
  run "sched3 16384 reserve2 $pct%" ATLB200_LU_RESERVE2_PCT=$pct -- $PROD 16384 5
done
run "sched3 16384 nb=384" ATLB200_LU_NB=384 -- $PROD 16384 5
run "sched3 16384 nb=256" ATLB200_LU_NB=256 -- $PROD 16384 5
run "sched3 32768" -- $PROD 32768 3
run "sched3 phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
run "potrf phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2 potrf
run "potrf 16384" -- $PROD 16384 10 potrf
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -40
echo DONE
