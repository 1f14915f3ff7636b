#!/bin/bash
# the driver's scaling launch at N = 8: default bench (incl. the in-bench training set-up on every rank)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2w_bench_n8.json 2> gpurun_out/r2w_bench_n8.err; tail -c 500 gpurun_out/r2w_bench_n8.json; echo
