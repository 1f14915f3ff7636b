#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest.log
timeout 300 python tools/variant_check.py > gpurun_out/r2s_variant.log 2>&1; echo "variant rc=$?"; grep -c bitwise-equal gpurun_out/r2s_variant.log; grep -i "different\|error\|Traceback" gpurun_out/r2s_variant.log | head
timeout 600 python tools/conv_lab.py --masks 8194 --full 8194 --rounds 3 --out gpurun_out/r2s_lab.json > gpurun_out/r2s_lab.log 2>&1; tail -4 gpurun_out/r2s_lab.log
