#!/bin/bash
# full ncu capture (with source) of one launch of each hot kernel; usage: gpu_ncu3.sh [quiva-gb] [fasta-gb] [kernel-regex]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KRE=${3:-k_qv_decode5|k_qv_code|k_qv_hist|k_fa_pack2|k_unpack2|k_qv_assemble|k_fa_measure2}
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KRE" \
    -o gpurun_out/prof -f python scripts/ncu_once.py ${1:-0.25} ${2:-0.25} > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out | head
