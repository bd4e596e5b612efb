#!/bin/bash
# development loop on the GPU box: parity tests (stop at first failure), one bench, and a small
# run with the decoder's debug counters.  args: bench size in GB (default 2), extra bench args
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
make -s -C tools > /dev/null 2>&1
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 1500 python bench.py --size-gb ${1:-2} --no-cpu $2 $3 $4 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?" ; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print("value",round(d['value'],2),"GB/s  ms/step",round(d['ms_per_step'],2),"e2e",round(d['e2e']['value'],2))
    print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['extra'].items() if k!='kernels_ms_per_step'})
    print(d['extra']['kernels_ms_per_step'])
    print(d['roofline']); print(d['path_roofline'])
except Exception as e:
    print("bench parse failed",e)
PY
tail -15 gpurun_out/bench.err
echo "== debug counters (0.25 GB)"
DEXB200_DEBUG=1 timeout 600 python bench.py --size-gb 0.25 --steps 1 --warmup 3 --no-cpu --no-extras 2>&1 >/dev/null | grep debug | tail -12
