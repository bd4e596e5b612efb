#!/bin/bash
# ncu evidence: launch list of a short bench run + one full capture of the named kernels.
# usage: gpu_ncu.sh <size-gb> <kernel-regex> [skip] [count]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SIZE=${1:-0.5}
KRE=${2:-k_qv_}
SKIP=${3:-0}
CNT=${4:-6}
CMD="python bench.py --size-gb $SIZE --steps 1 --warmup 3 --no-extras --no-cpu"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1 ; echo "rc=$?"
echo "== full capture of $KRE"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT \
    -o gpurun_out/prof -f $CMD > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -20
