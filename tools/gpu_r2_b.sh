#!/bin/bash
# round-2 call b: new pinned GPU tests + headline tests, conv lab (pair kernel per layer), quick bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 900 python tools/conv_lab.py --out gpurun_out/r2b_lab.json > gpurun_out/r2b_lab.log 2>&1
tail -30 gpurun_out/r2b_lab.log
