#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/r2last2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2last2_pytest.log
timeout 300 python tools/train_bench.py --batch 4 --steps 10 --warmup 3 > gpurun_out/r2last2_train.json 2>/dev/null; cut -c1-200 gpurun_out/r2last2_train.json
