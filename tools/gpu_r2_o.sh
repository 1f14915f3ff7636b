#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1_stack" -s 4 -c 2 -o gpurun_out/r2o_prof_stack python tools/conv_lab.py --masks 9218 --rounds 1 --full "" --out gpurun_out/r2o_lab3.json > gpurun_out/r2o_ncu2.log 2>&1
ls -la gpurun_out/r2o_prof_*; tail -3 gpurun_out/r2o_ncu2.log
