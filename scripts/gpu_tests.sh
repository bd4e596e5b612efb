#!/bin/bash
# the whole GPU suite (optionally -k expr)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x ${1:+-k "$1"} --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
