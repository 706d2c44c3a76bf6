#!/bin/bash
# Third matrix: cost of the fence / of -dlcm=cg, look-ahead traces, then the whole GPU suite.
cd "$(dirname "$0")/.."
BIN=tools/bin/lu_race
PROD=math_atlas_b200/libatlas_b200.so
FORCED="ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_RESERVE=0 ATLB200_GEMM_CTA_CAP=1000000"
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $label :: ${envs[*]} :: $*"
  env "${envs[@]}" timeout 900 $BIN "$@" 2>&1 | tail -n 80
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in prod nofence l1; do
  L=tools/bin/libatlas_b200_$v.so; [ $v = prod ] && L=$PROD
  run "gemm $v 16384" -- $L 16384 5 gemm
  run "gemm $v 8192" -- $L 8192 8 gemm
  run "gemm $v 4096" -- $L 4096 10 gemm
  run "gemm $v 2048" -- $L 2048 10 gemm
  run "forced race $v 8192" $FORCED ATLB200_LU_PW=0 -- $L 8192 60
  run "getrf $v 16384" -- $L 16384 8
  run "getrf $v 8192" -- $L 8192 8
  run "potrf $v 16384" -- $L 16384 8 potrf
done
run "forced race nofence 16384" $FORCED ATLB200_LU_PW=0 -- tools/bin/libatlas_b200_nofence.so 16384 30
run "trace 16384 pw512" ATLB200_LU_TRACE=1 -- $PROD 16384 2
run "trace 16384 pw128" ATLB200_LU_TRACE=1 ATLB200_LU_PW=128 -- $PROD 16384 2
run "trace 8192 pw512" ATLB200_LU_TRACE=1 -- $PROD 8192 2
echo "=== pytest -m gpu (whole suite)"
timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 2>&1 | tail -40
echo DONE
