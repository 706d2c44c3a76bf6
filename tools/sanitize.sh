#!/bin/bash
# compute-sanitizer over the round-2 kernels (small instances, tools/sanitize_cases.py); output -> gpurun_out/sanitize_*.log
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  for c in gemm lu qr; do
    echo "=== compute-sanitizer --tool $tool : $c"
    timeout 240 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_cases.py $c 2>&1 | grep -v "^=========$" | tail -n 8
  done
done
