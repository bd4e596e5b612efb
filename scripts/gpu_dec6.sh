#!/bin/bash
# lane decoder: tests, the decoders side by side, optionally one ncu --set full capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SIZE=${1:-0.5}
timeout 600 python -m pytest tests -q -m gpu -x -k "lane or route" 2>&1 | tail -4
timeout 600 python scripts/dec_probe.py $SIZE 2>&1 | tail -12
if [ "$2" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qv_decode6 -s 2 -c 1 \
    -o gpurun_out/dec6 -f python scripts/dec_probe.py ${3:-0.25} decoder=6 > gpurun_out/ncu_dec6.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_metrics.py gpurun_out/dec6.ncu-rep 2>&1 | tail -24
python scripts/ncu_stalls.py gpurun_out/dec6.ncu-rep 2>&1 | tail -3
fi
