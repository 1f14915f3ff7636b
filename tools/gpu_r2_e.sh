#!/bin/bash
# round-2 call e: 8-warp epilogue of the rolling conv1 kernel, new defaults (roll + conv2 pair), graph fix; tests, lab, bench, training convergence probe
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2e_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2e_variant.log
tail -5 gpurun_out/r2e_variant.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
timeout 900 python tools/conv_lab.py --rounds 2 --masks 1026,0,1024,2 --full 1026,0,1024,2 --out gpurun_out/r2e_lab.json > gpurun_out/r2e_lab.log 2>&1
tail -10 gpurun_out/r2e_lab.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 600 gpurun_out/r2e_bench.json
timeout 600 python tools/train_synth.py --steps 200 --batch 16 --lr 1e-4 --eval-every 50 --out gpurun_out/r2e_train_lr1e-4.json > gpurun_out/r2e_train_a.log 2>&1
tail -6 gpurun_out/r2e_train_a.log
timeout 600 python tools/train_synth.py --steps 200 --batch 16 --lr 1e-3 --eval-every 50 --out gpurun_out/r2e_train_lr1e-3.json > gpurun_out/r2e_train_b.log 2>&1
tail -6 gpurun_out/r2e_train_b.log
