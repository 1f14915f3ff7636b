#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 240 python tools/variant_check.py > gpurun_out/r2l_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2l_variant.log
grep -c "bitwise-equal" gpurun_out/r2l_variant.log; grep -v "bitwise-equal" gpurun_out/r2l_variant.log | tail -8
timeout 400 python tools/conv_lab.py --rounds 3 --masks 1026,9218,0 --full 1026,9218,0 --out gpurun_out/r2l_lab.json > gpurun_out/r2l_lab.log 2>&1
tail -9 gpurun_out/r2l_lab.log | cut -c1-150
