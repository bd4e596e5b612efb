#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run, DRAM traffic per kernel on the full
# 2 GB workload, and one full capture of the heaviest kernels.  args: size-gb (default 2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SIZE=${1:-2}
CMD="python bench.py --size-gb $SIZE --steps 2 --warmup 3 --no-extras --no-cpu"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1 ; echo "rc=$?"
echo "== dram traffic"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_ -s 60 -c 40 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1 ; echo "rc=$?"
echo "== full capture"
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_qv_decode5|k_qv_code|k_qv_compact|k_qv_hist|k_pred_slots" \
    -s 24 -c 8 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -20
echo "== 2-bit kernels (1 GB fasta), full capture"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_fa_pack3|k_unpack3|k_pred_slots|k_fa_measure2" \
    -o gpurun_out/prof_pack -f python scripts/ncu_once.py 0.05 1.0 > gpurun_out/ncu_pack.log 2>&1 ; echo "rc=$?"
