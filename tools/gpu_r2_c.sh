#!/bin/bash
# round-2 call c: kernel-variant equivalence (rolling conv1, CTA pairs, CUDA graph), tests, lab
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2c_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2c_variant.log
tail -40 gpurun_out/r2c_variant.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 900 python tools/conv_lab.py --rounds 3 --out gpurun_out/r2c_lab.json > gpurun_out/r2c_lab.log 2>&1
tail -30 gpurun_out/r2c_lab.log
