#!/bin/bash
# full ncu capture of named kernels during prof_paths.py (small size); usage: gpu_ncu2.sh <size-gb> <kernel-regex> [count] [fasta-gb]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SIZE=${1:-0.25}; KRE=${2:-k_qv_}; CNT=${3:-8}; FA=${4:-0}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$KRE -c $CNT \
    -o gpurun_out/prof -f python scripts/prof_paths.py $SIZE $FA > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out | head
