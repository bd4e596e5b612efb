#!/bin/bash
# a kernel change: the tests named by $1 (-k expression), then a short bench line with the kernel table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "$1" 2>&1 | tail -3
timeout 600 python bench.py --no-extras --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['path_roofline'].items()})
print(d['extra']['kernels_ms_per_step'])
PY
