#!/bin/bash
# round-2 call h: slim MMA loop + linear epilogue of the rolling conv1 kernel; bench with trained weights; all GPU tests
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2h_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2h_variant.log
tail -4 gpurun_out/r2h_variant.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
timeout 600 python tools/conv_lab.py --rounds 2 --masks 1026,0 --full 1026,0 --out gpurun_out/r2h_lab.json > gpurun_out/r2h_lab.log 2>&1
tail -5 gpurun_out/r2h_lab.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 2500 gpurun_out/r2h_bench.json; tail -5 gpurun_out/r2h_bench.err
