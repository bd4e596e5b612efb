#!/bin/bash
# the TMA (cp.async.bulk) experiment on the newline index: correctness through the route test, timing
# side by side, and ncu's key numbers for both kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_paths.py -q -m gpu -x -k "index_bulk or shard" 2>&1 | tail -3
timeout 600 python scripts/tma_probe.py 2 2>&1 | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:k_pred_slots -s 4 -c 12 --csv --log-file gpurun_out/tma_ncu.csv python scripts/tma_probe.py 2 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/tma_ncu.csv')) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    try: agg[r[ix['Kernel Name']].split('(')[0][-28:] + ' grid ' + r[ix['Grid Size']]][r[ix['Metric Name']]].append(float(r[ix['Metric Value']].replace(',', '')))
    except Exception: pass
for k, m in agg.items():
    print(k, {a: round(sum(v) / len(v), 3) for a, v in m.items()})
PY
