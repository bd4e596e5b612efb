#!/bin/bash
# N-GPU bench exactly as the driver launches it (no extra flags).  usage: gpu_multi.sh N [extra bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
timeout 860 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 0 > gpurun_out/ref_n$N.json 2> gpurun_out/ref_n$N.err
echo "reference N=$N rc=$?"; tail -c 300 gpurun_out/ref_n$N.json
