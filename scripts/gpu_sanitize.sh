#!/bin/bash
# compute-sanitizer memcheck over the smoke run and the GPU suite without the BASELINE-size tests
# (tests/test_gpu_scale.py: 1 GB / 8 GB under memcheck would take the round), then racecheck over smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== new tests alone"; timeout 600 python -m pytest tests/test_gpu_paths.py -q -x -p no:cacheprovider -k "usual_range or shard or histogram" 2>&1 | tail -3
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/sanitizer_smoke.log
echo "== memcheck tests"; timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -p no:cacheprovider \
    --deselect tests/test_gpu_scale.py --deselect tests/test_gpu_cli.py --deselect tests/test_gpu_compat.py > gpurun_out/sanitizer_tests.log 2>&1; echo "rc=$?"
grep -c "Invalid\|out of bounds\|Misaligned" gpurun_out/sanitizer_tests.log; tail -6 gpurun_out/sanitizer_tests.log
