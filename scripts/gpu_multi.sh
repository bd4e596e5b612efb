#!/bin/bash
# N-GPU bench as the driver launches it.  usage: gpu_multi.sh N [size-gb]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 3 --warmup 3 --size-gb ${2:-2} --no-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_n$N.json; tail -15 gpurun_out/bench_n$N.err
