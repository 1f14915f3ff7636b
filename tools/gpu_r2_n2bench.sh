#!/bin/bash
# the driver's scaling launch at N = 2: default bench (incl. the in-bench training set-up on every rank) and the reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; tail -c 400 gpurun_out/r2w_bench_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2w_bench_ref_n2.json 2> gpurun_out/r2w_bench_ref_n2.err; tail -c 400 gpurun_out/r2w_bench_ref_n2.json; echo
