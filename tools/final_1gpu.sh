#!/bin/bash
# Round-2 closing run on one GPU: the whole -m gpu suite, smoke(), the bench line of both arms, the sweeps, and the profiler
# captures the numbers in DESIGN.md / profiles/ come from (a number printed under ncu is never a bench value).
cd "$(dirname "$0")/.."
PROD=math_atlas_b200/libatlas_b200.so
(time python -m pytest tests -m gpu -q) > gpurun_out/final_gputests.txt 2>&1; tail -4 gpurun_out/final_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; tail -1 gpurun_out/final_smoke.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 3000 gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cat gpurun_out/final_bench_ref.json
python tools/gemm_small.py > gpurun_out/final_gemm_small.jsonl 2>&1
python tools/sweep.py > gpurun_out/final_sweep.jsonl 2> gpurun_out/final_sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-big > gpurun_out/final_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_kernel -s 1 -c 1 -o gpurun_out/final_dgemm \
    tools/bin/lu_race $PROD 16384 1 gemm > gpurun_out/final_ncu_d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final_getrf16k_launches.csv \
    tools/bin/lu_race $PROD 16384 1 > gpurun_out/final_getrf16k.log 2>&1
ncu --set full --clock-control none -k regex:laswp_par_kernel -s 40 -c 2 -o gpurun_out/final_laswp \
    tools/bin/lu_race $PROD 16384 1 > gpurun_out/final_ncu_l.log 2>&1
ls -la gpurun_out/final_*
