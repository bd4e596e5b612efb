#!/bin/bash
# direct placement of discovered entries: whole suite, bench, host phases; "ncu" as $1 adds one full
# capture of the heaviest kernels on the 2 GB bench file (with source)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py --no-cpu > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_direct.json
DEXB200_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ph.json 2> gpurun_out/ph.err; echo rc=$?
grep "host phases\|candidates" gpurun_out/ph.err | tail -8
if [ "$1" = "ncu" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_qv_decode5|k_qv_code|k_qv_hist_run" \
    -s 12 -c 4 -o gpurun_out/prof_r2 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_full.log 2>&1 ; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
fi
