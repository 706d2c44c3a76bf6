#!/bin/bash
# Round-2 closing run on 8 GPUs of one box: the driver-style 8-rank bench line, the distributed-LU stress (50 runs, pivot hash +
# factor checksum), DPOTRF, and the one-process multi-device entry points (tools/multi_dev_bench.py).
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/final_bench_${N}gpu.json 2> gpurun_out/final_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 2600 gpurun_out/final_bench_${N}gpu.json
timeout 600 $TR --master-port 29542 tools/dist_bench.py --size 16384 --what getrf --reps 50 --check 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
timeout 600 $TR --master-port 29543 tools/dist_bench.py --size 16384 --what potrf --reps 20 --check 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
timeout 600 $TR --master-port 29544 tools/dist_bench.py --size 65536 --what getrf --reps 4 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
timeout 900 python tools/multi_dev_bench.py 16384 gemm,getrf,potrf
timeout 900 python tools/multi_dev_bench.py 32768 getrf,potrf
