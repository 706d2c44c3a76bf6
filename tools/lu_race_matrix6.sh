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
run "gemm 16384" -- $PROD 16384 6 gemm
for n in 256 512 1024 2000 2048 4096; do
  run "gemm $n auto" -- $PROD $n 20 gemm
  run "gemm $n bn=128" ATLB200_GEMM_BN=128 -- $PROD $n 20 gemm
  run "gemm $n bn=64" ATLB200_GEMM_BN=64 -- $PROD $n 20 gemm
done
echo "=== gemm selftest BN=64"
ATLB200_GEMM_BN=64 timeout 600 tools/bin/gemm_selftest check 2>&1 | tail -4
echo "=== gemm selftest auto"
timeout 600 tools/bin/gemm_selftest check 2>&1 | tail -2
run "forced race 8192" $FORCED -- $PROD 8192 40
run "forced diag VERIFY 8192" $FORCED ATLB200_DIAG_VERIFY=1 -- $DIAG 8192 8
run "default determinism 16384" -- $PROD 16384 30
run "default determinism 8192" -- $PROD 8192 40
run "default determinism 4096" -- $PROD 4096 40
for n in 16384 8192 4096; do
  run "getrf $n aux=0" ATLB200_LU_AUX=0 -- $PROD $n 5
  run "getrf $n defer_left=0" ATLB200_LU_DEFER_LEFT=0 -- $PROD $n 5
  run "getrf $n aux=0 defer_left=0" ATLB200_LU_AUX=0 ATLB200_LU_DEFER_LEFT=0 -- $PROD $n 5
  run "getrf $n nb=384" ATLB200_LU_NB=384 -- $PROD $n 5
done
run "phase trace 16384" ATLB200_LU_TRACE=2 -- $PROD 16384 2
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -40
echo DONE
