#!/bin/bash
# two GPUs: the tests named by $1 on GPU 0, then the driver's exact N=2 command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "$1" 2>&1 | tail -3
bash scripts/gpu_multi.sh 2 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | cut -c1-600
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['path_roofline'].items()}, d['e2e']['value'])
PY
