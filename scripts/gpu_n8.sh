#!/bin/bash
# 8-GPU session: the driver's N=8 bench command, then dexqv_mg on a 64 GB file (cfg4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_legacy_formats.py -q -m gpu 2>&1 | tail -2
bash scripts/gpu_multi.sh 8 2>&1 | cut -c1-600
bash scripts/gpu_mg.sh 8 ${1:-64} 4 2>&1 | cut -c1-2500
