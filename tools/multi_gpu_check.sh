#!/bin/bash
# usage: tools/multi_gpu_check.sh N   (under gpurun --gpus N): multi-device tests + the N-rank bench line
N=${1:-2}
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -15
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_r02_${N}gpu.json; tail -8 gpurun_out/bench_r02_${N}gpu.err
