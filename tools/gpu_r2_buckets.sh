#!/bin/bash
# bucket-size sweep of the overlapped gradient all-reduce.  usage: gpurun --gpus N -- 'bash tools/gpu_r2_buckets.sh N'
N=${1:-4}
set -x
mkdir -p gpurun_out
PORT=29600
for mb in 128 64 512; do
  PORT=$((PORT+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT tools/train_bench.py --batch 4 --steps 10 --warmup 3 --bucket-mb $mb > gpurun_out/r2b_train_n${N}_mb${mb}.json 2> gpurun_out/r2b_train_n${N}_mb${mb}.err
  tail -1 gpurun_out/r2b_train_n${N}_mb${mb}.json | cut -c1-330
done
