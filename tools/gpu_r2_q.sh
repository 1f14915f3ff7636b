#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
timeout 600 python tools/conv_lab.py --masks 8194 --full 8194 --rounds 3 --out gpurun_out/r2q_lab.json > gpurun_out/r2q_lab.log 2>&1; tail -4 gpurun_out/r2q_lab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zoom_fused|raster_resolve|pack_obs4" -s 20 -c 12 --csv --log-file gpurun_out/r2q_zoom.csv python bench.py --steps 1 --warmup 3 --step-batches 4 --slots 1 --no-cpu-baseline --no-fast-mode --train-steps 0 > gpurun_out/r2q_ncu_bench.log 2>&1
grep -o '"zoom_fused[^,]*,[^,]*,[^,]*\|[0-9.]*"$' gpurun_out/r2q_zoom.csv | tail -12
