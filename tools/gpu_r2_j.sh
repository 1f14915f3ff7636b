#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2j_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2j_variant.log
grep -c "bitwise-equal" gpurun_out/r2j_variant.log; grep -v "bitwise-equal" gpurun_out/r2j_variant.log | tail -5
timeout 600 python tools/conv_lab.py --rounds 3 --masks 1026,5122,0 --full 1026,5122,0 --out gpurun_out/r2j_lab.json > gpurun_out/r2j_lab.log 2>&1
tail -9 gpurun_out/r2j_lab.log
