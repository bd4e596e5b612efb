#!/bin/bash
# quick iteration: parity tests (-x), then per-path kernel tables.  args: quiva GB, fasta GB
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
make -s -C tools > /dev/null 2>&1
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== paths" ; DEXB200_DEBUG=${DEBUG:-} timeout 900 python scripts/prof_paths.py ${1:-2} ${2:-1} > gpurun_out/paths.log 2>&1 ; echo "rc=$?"; grep -v "v5 table" gpurun_out/paths.log | tail -80
