#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2u_variant.log 2>&1; echo "variant rc=$?"; grep -c bitwise-equal gpurun_out/r2u_variant.log; grep -i "different\|error\|Traceback" gpurun_out/r2u_variant.log | head
timeout 600 python tools/conv_lab.py --masks 8194 --full 8194 --rounds 3 --out gpurun_out/r2u_lab.json > gpurun_out/r2u_lab.log 2>&1; tail -4 gpurun_out/r2u_lab.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1_stack" -s 4 -c 1 -o gpurun_out/r2u_prof_stack python tools/conv_lab.py --masks 8194 --rounds 1 --full "" --out gpurun_out/r2u_lab3.json > gpurun_out/r2u_ncu.log 2>&1
tail -2 gpurun_out/r2u_ncu.log
