#!/bin/bash
# round-end evidence: parity tests, the default bench line, and the ncu passes behind profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?" ; tail -c 1500 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
bash scripts/gpu_profiles.sh 2
