#!/bin/bash
# Round-2 profiler captures (1 GPU): launch list of the bench command, ncu --set full of the dominant DGEMM kernel and of the
# default s/c tensor-core kernel (sgemm_tc2_kernel), launch list of one DGETRF N=16384.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
PROD=math_atlas_b200/libatlas_b200.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-big > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_kernel -s 1 -c 1 -o gpurun_out/r02_dgemm \
    tools/bin/lu_race $PROD 16384 1 gemm > gpurun_out/r02_ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sgemm_tc2_kernel -s 1 -c 1 -o gpurun_out/r02_sgemm_tc2 \
    python tools/prof_gemm.py s 8192 > gpurun_out/r02_ncu_s.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_getrf16k_launches.csv \
    tools/bin/lu_race $PROD 16384 1 > gpurun_out/r02_getrf16k.log 2>&1
ls -la gpurun_out/r02_*
