#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export DIM_CONV1_DBG_ALIGN=1; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"conv1_stack" -s 4 -c 3 --csv --log-file gpurun_out/r2v_align$v.csv python tools/conv_lab.py --masks 8194 --rounds 1 --full "" --out gpurun_out/r2v_lab$v.json > gpurun_out/r2v_ncu$v.log 2>&1
  grep -o '"conv1_stack[^"]*".*' gpurun_out/r2v_align$v.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -9
done
