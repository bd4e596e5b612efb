#!/bin/bash
# First GPU call of the next round: everything round 1 could not re-measure after its GPU minutes ran out.
#   1 GPU:  gpurun --timeout 900 -- scripts/gpu_next_round_first.sh
#   2 GPUs: gpurun --gpus 2 --timeout 600 -- scripts/gpu_next_round_first.sh 2
# Writes gpurun_out/next_*.{log,json,err}.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  # (a) the whole GPU suite, including the tests added at the end of round 1 (fuzz: ran green once;
  #     test_gpu_zz_interactive: never run on a GPU yet)
  timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/next_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/next_pytest.log
  # (b) cfg4 shard sizes: 8 GB per GPU (the 8-GPU share of a 64 GB file); 32 GB is the 2-GPU share
  for G in 8; do
    timeout 600 python bench.py --steps 3 --warmup 3 --size-gb $G --no-extras --no-cpu \
        > gpurun_out/next_bench_${G}gb.json 2> gpurun_out/next_bench_${G}gb.err; echo "bench ${G} GB rc=$?"
    tail -c 600 gpurun_out/next_bench_${G}gb.json
  done
else
  # (c) exit status of an N>1 run after the teardown fix (every rank used to end in SIGABRT)
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-extras > gpurun_out/next_bench_n$N.json 2> gpurun_out/next_bench_n$N.err
  echo "torchrun rc=$? (must be 0)"; grep -c "context is destroyed" gpurun_out/next_bench_n$N.err
  tail -c 400 gpurun_out/next_bench_n$N.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
      bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/next_ref_n$N.json 2> gpurun_out/next_ref_n$N.err
  echo "reference arm rc=$?"; wc -l gpurun_out/next_ref_n$N.json
fi
