#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2r_variant.log 2>&1; echo "variant rc=$?"; grep -c bitwise-equal gpurun_out/r2r_variant.log; grep -i "different\|error\|Traceback" gpurun_out/r2r_variant.log | head
timeout 600 python tools/conv_lab.py --masks 8194 --full 8194 --rounds 3 --out gpurun_out/r2r_lab.json > gpurun_out/r2r_lab.log 2>&1; tail -4 gpurun_out/r2r_lab.log
