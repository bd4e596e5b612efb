#!/bin/bash
# dexqv_mg on the N GPUs of the box.  usage: gpu_mg.sh N [big_gb] [ref_max_gb]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; df -h /dev/shm | tail -1; free -g | head -2
timeout 3000 python scripts/mg_check.py ${1:-2} ${2:-0.5} ${3:-4} > gpurun_out/mg_n${1:-2}.jsonl 2> gpurun_out/mg_n${1:-2}.err
echo "rc=$?"; cut -c1-1500 gpurun_out/mg_n${1:-2}.jsonl; tail -5 gpurun_out/mg_n${1:-2}.err
