#!/bin/bash
# development loop: targeted tests, then a short bench.   usage: gpu_dev.sh "<pytest -k expr>" [bench args...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=${1:-lane}; shift
timeout 900 python -m pytest tests -q -m gpu -x -k "$K" > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/dev_pytest.log
if [ "$1" != "nobench" ]; then
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/dev_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/dev_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
    print(d['roofline'])
    print({k:v['ms'] for k,v in d['path_roofline'].items()})
    print(d['extra']['kernels_ms_per_step'])
    print({k:v for k,v in d['extra'].items() if k in ('dexta_gbs','undexta_gbs','dexar_gbs','undexar_gbs','length_sweep')})
except Exception as e: print('no bench line', e)
PY
fi
