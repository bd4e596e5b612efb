#!/bin/bash
# tests + bench on the GPU box; args: size-gb for a quick first bench (optional)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_gpu.log
if [ -n "$1" ]; then
  echo "== quick bench $1 GB" ; timeout 600 python bench.py --size-gb $1 --steps 2 --warmup 3 --cpu-sample-mb 8 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err ; echo "rc=$?" ; tail -c 3000 gpurun_out/bench_quick.json ; tail -20 gpurun_out/bench_quick.err
fi
echo "== bench" ; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?" ; cat gpurun_out/bench.json ; tail -20 gpurun_out/bench.err
