#!/bin/bash
# the whole GPU suite (no -x: every failure is listed), then the default bench line and the ncu launch
# list of the same command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log | cut -c1-260
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cut -c1-1800 gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_launch.log 2>&1 ; echo "ncu rc=$?"
