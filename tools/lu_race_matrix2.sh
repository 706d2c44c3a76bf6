#!/bin/bash
# Second matrix: after the stage-release fix in the DGEMM/ZGEMM kernels (mbar_release).
cd "$(dirname "$0")/.."
BIN=tools/bin/lu_race
PROD=math_atlas_b200/libatlas_b200.so
DIAG=tools/bin/libatlas_b200_diag.so
FORCED="ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_RESERVE=0 ATLB200_GEMM_CTA_CAP=1000000"
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $label :: ${envs[*]} :: $*"
  env "${envs[@]}" timeout 900 $BIN "$@" 2>&1 | tail -n 30
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for pw in 0 128; do
run "forced 8192 pw=$pw" $FORCED ATLB200_LU_PW=$pw -- $PROD 8192 60
run "forced 16384 pw=$pw" $FORCED ATLB200_LU_PW=$pw -- $PROD 16384 40
done
run "forced diag VERIFY 8192 pw=0" $FORCED ATLB200_LU_PW=0 ATLB200_DIAG_VERIFY=1 -- $DIAG 8192 20
run "overlap persistent reserve=0 16384" ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_RESERVE=0 -- $PROD 16384 60
for pw in 0 64 128 256 512; do
run "overlap default-reserve 16384 pw=$pw" ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_PW=$pw -- $PROD 16384 12
run "no-overlap 16384 pw=$pw" ATLB200_LU_PW=$pw -- $PROD 16384 6
run "overlap 8192 pw=$pw" ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_PW=$pw -- $PROD 8192 12
run "overlap 4096 pw=$pw" ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_PW=$pw -- $PROD 4096 12
run "2048 pw=$pw" ATLB200_LU_PW=$pw -- $PROD 2048 12
run "1024 pw=$pw" ATLB200_LU_PW=$pw -- $PROD 1024 12
done
run "potrf 16384" -- $PROD 16384 20 potrf
run "potrf 8192" -- $PROD 8192 20 potrf
echo "=== gemm selftest"
timeout 600 tools/bin/gemm_selftest 2>&1 | tail -40
echo "=== pytest getrf subset"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "getrf or lu or potrf" 2>&1 | tail -15
echo DONE
