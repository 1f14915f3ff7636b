#!/bin/bash
# round-2 call d: warp-uniform elected MMA/TMA issue -- equivalence of variants, full GPU tests, lab, bench
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_check.py > gpurun_out/r2d_variant.log 2>&1; echo "variant rc=$?" >> gpurun_out/r2d_variant.log
tail -32 gpurun_out/r2d_variant.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
timeout 900 python tools/conv_lab.py --rounds 2 --out gpurun_out/r2d_lab.json > gpurun_out/r2d_lab.log 2>&1
tail -14 gpurun_out/r2d_lab.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 1500 gpurun_out/r2d_bench.json
