#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 240 python tools/variant_check.py > gpurun_out/r2n_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2n_variant.log
grep -c "bitwise-equal" gpurun_out/r2n_variant.log; grep -v "bitwise-equal" gpurun_out/r2n_variant.log | tail -4
timeout 400 python tools/conv_lab.py --rounds 3 --masks 1026,9218,0 --full 1026,9218,0 --out gpurun_out/r2n_lab.json > gpurun_out/r2n_lab.log 2>&1
tail -9 gpurun_out/r2n_lab.log | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1_" -s 44 -c 2 -o gpurun_out/r2n_prof_roll python tools/conv_lab.py --masks 1026 --rounds 1 --full "" --out gpurun_out/r2n_lab2.json > gpurun_out/r2n_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1_" -s 44 -c 2 -o gpurun_out/r2n_prof_stack python tools/conv_lab.py --masks 9218 --rounds 1 --full "" --out gpurun_out/r2n_lab3.json > gpurun_out/r2n_ncu2.log 2>&1
ls -la gpurun_out/r2n_prof_*
