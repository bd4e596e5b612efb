#!/bin/bash
# everything profiles/ is made of (scripts/make_profiles.py rNN afterwards, here): the whole GPU suite,
# the default bench line, the ncu launch list / DRAM traffic / full captures of scripts/gpu_profiles.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/bench_ref.json
bash scripts/gpu_profiles.sh 2 2>&1 | tail -12
