#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2last3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2last3_pytest.log
