#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the smoke run and the scan / decode route tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/racecheck_smoke.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_paths.py -q -p no:cacheprovider \
    -k "histogram or shard or usual_range or route-default or no_direct" > gpurun_out/racecheck_paths.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/racecheck_paths.log
