#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DEXB200_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ph.json 2> gpurun_out/ph.err; echo rc=$?
grep "host phases" gpurun_out/ph.err | tail -8
