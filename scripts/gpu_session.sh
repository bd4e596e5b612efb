#!/bin/bash
# One GPU-box session: smoke, parity tests, (optionally) sanitizer + bench.  Everything is logged
# under gpurun_out/ so it can be read back in the build container.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
if [ "$1" == "sanitize" ]; then
  echo "== compute-sanitizer" ; timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1 ; echo "sanitizer rc=$?" ; tail -30 gpurun_out/sanitizer.log
fi
