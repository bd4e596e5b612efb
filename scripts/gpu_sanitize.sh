#!/bin/bash
# compute-sanitizer memcheck over the smoke run and the 2-bit / route parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== new test alone"; timeout 600 python -m pytest tests/test_gpu_paths.py -q -x -p no:cacheprovider -k "lattice" 2>&1 | tail -3
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/sanitizer_smoke.log
echo "== memcheck tests"; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_paths.py -q -x -p no:cacheprovider -k "lattice or scratch or default" > gpurun_out/sanitizer_tests.log 2>&1; echo "rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/sanitizer_tests.log; tail -6 gpurun_out/sanitizer_tests.log
