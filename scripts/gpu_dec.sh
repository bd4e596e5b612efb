#!/bin/bash
# decoder study: debug counters + full ncu capture of the decode kernel; usage: gpu_dec.sh [quiva-gb]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DEXB200_DEBUG_DEC=1 timeout 600 python bench.py --size-gb ${1:-0.25} --steps 1 --warmup 3 --no-cpu --no-extras 2>&1 >/dev/null | grep "v5 table" | tail -8
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_qv_decode5" \
    -o gpurun_out/prof -f python scripts/ncu_once.py ${1:-0.25} 0.05 > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
tail -3 gpurun_out/ncu_full.log
