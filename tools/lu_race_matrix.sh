#!/bin/bash
# Experiment matrix for the LU look-ahead nondeterminism.  Run on a GPU box:
#   bash tools/lu_race_matrix.sh > gpurun_out/race.log 2>&1
cd "$(dirname "$0")/.."
BIN=tools/bin/lu_race
PROD=math_atlas_b200/libatlas_b200.so
DIAG=tools/bin/libatlas_b200_diag.so
FORCED="ATLB200_LU_OVERLAP_GEMM=1 ATLB200_LU_RESERVE=0 ATLB200_GEMM_CTA_CAP=1000000"
run() {  # label, env..., -- args
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $label :: ${envs[*]} :: $*"
  env "${envs[@]}" timeout 600 $BIN "$@" 2>&1 | tail -n 60
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
run "default product 8192" -- $PROD 8192 30
run "default product 16384" -- $PROD 16384 20
run "forced product 4096 nb256" $FORCED ATLB200_LU_NB=256 -- $PROD 4096 40
run "forced product 8192" $FORCED -- $PROD 8192 30
run "forced product 16384" $FORCED -- $PROD 16384 20
run "forced diag 8192 (no verify)" $FORCED -- $DIAG 8192 30
run "forced diag 8192 VERIFY" $FORCED ATLB200_DIAG_VERIFY=1 -- $DIAG 8192 10
run "forced diag 16384 VERIFY" $FORCED ATLB200_DIAG_VERIFY=1 -- $DIAG 16384 6
run "detector sanity: side stream empty, VERIFY" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=15 -- $DIAG 8192 6
run "USYNC before update GEMM" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_USYNC=1 -- $DIAG 8192 10
run "global-barrier leaf" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_NOCLUSTER=1 -- $DIAG 8192 10
run "leaf -> nothing" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=1 -- $DIAG 8192 10
run "leaf -> cluster spin" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=16 -- $DIAG 8192 10
run "leaf -> plain spin" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=32 -- $DIAG 8192 10
run "side laswp -> nothing" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=2 -- $DIAG 8192 10
run "side trsm -> nothing" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=4 -- $DIAG 8192 10
run "side gemm -> nothing" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=8 -- $DIAG 8192 10
run "only leaf on side (cluster)" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=14 -- $DIAG 8192 10
run "only cluster spin on side" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=30 -- $DIAG 8192 10
run "only plain spin on side" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=46 -- $DIAG 8192 10
run "victim on generic kernel" $FORCED ATLB200_DIAG_VERIFY=1 ATLB200_DIAG_PNOP=64 -- $DIAG 4096 10
echo "=== compute-sanitizer memcheck (forced, N=4096 NB=256)"
env $FORCED ATLB200_LU_NB=256 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 $BIN $PROD 4096 2 2>&1 | tail -n 40
echo "=== compute-sanitizer racecheck (forced, N=2048 NB=256)"
env $FORCED ATLB200_LU_NB=256 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 $BIN $PROD 2048 2 2>&1 | tail -n 40
echo "=== compute-sanitizer initcheck (forced, N=2048 NB=256)"
env $FORCED ATLB200_LU_NB=256 timeout 900 compute-sanitizer --tool initcheck --print-limit 20 $BIN $PROD 2048 2 2>&1 | tail -n 40
echo DONE
