#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2p_variant.log 2>&1; echo "variant rc=$?"; grep -c bitwise-equal gpurun_out/r2p_variant.log; grep -i "different\|error\|Traceback" gpurun_out/r2p_variant.log | head
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2p_pytest.log
timeout 600 python tools/conv_lab.py --masks 8194,2 --full 8194,2 --rounds 2 --out gpurun_out/r2p_lab.json > gpurun_out/r2p_lab.log 2>&1; tail -6 gpurun_out/r2p_lab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 3 --step-batches 4 --slots 1 --no-cpu-baseline --no-fast-mode --train-steps 0 > gpurun_out/r2p_ncu_bench.log 2>&1
tail -2 gpurun_out/r2p_ncu_bench.log
