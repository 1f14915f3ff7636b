#!/bin/bash
# multi-GPU measurements of BASELINE.json configs C3 (64 instances, 8 per GPU at N = 8), C5 (128 instances strong-scaled) and C4
# (training step, per-GPU batch 4, NCCL gradient all-reduce).  usage: gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
N=${1:-1}
set -x
mkdir -p gpurun_out
PORT=29511
run() {  # run <tag> <args...>
  tag=$1; shift
  if [ "$N" = "1" ]; then
    timeout 600 python "$@" > gpurun_out/r2m_${tag}_n${N}.json 2> gpurun_out/r2m_${tag}_n${N}.err
  else
    PORT=$((PORT+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@" > gpurun_out/r2m_${tag}_n${N}.json 2> gpurun_out/r2m_${tag}_n${N}.err
  fi
  tail -c 400 gpurun_out/r2m_${tag}_n${N}.json; echo
}
run train tools/train_bench.py --batch 4 --steps 10 --warmup 3
run c5 bench.py --gpus $N --config c5 --scaling strong --batch $((128 / N)) --steps 6 --warmup 3 --step-batches 8 --no-cpu-baseline --no-fast-mode --train-steps 0
if [ "$N" = "8" ] || [ "$N" = "1" ]; then
  run c3 bench.py --gpus $N --config c3 --batch 8 --steps 6 --warmup 3 --step-batches 32 --no-cpu-baseline --no-fast-mode --train-steps 0
fi
if [ "$N" != "1" ]; then
  run allreduce tools/allreduce_bench.py
fi
