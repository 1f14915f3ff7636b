#!/bin/bash
# final evidence of round 2 on one B200: GPU tests, smoke, default bench + reference arm, launch list, full ncu capture of one
# iteration's kernels, parity report.  Outputs under gpurun_out/r2z_*.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; tail -c 600 gpurun_out/r2z_bench_ref.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 1 --warmup 3 --step-batches 4 --slots 1 --no-cpu-baseline --no-fast-mode --train-steps 0 > gpurun_out/r2z_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 130 -c 30 -o gpurun_out/r2z_prof python bench.py --steps 1 --warmup 1 --step-batches 2 --slots 1 --no-cpu-baseline --no-fast-mode --train-steps 0 > gpurun_out/r2z_ncu_full.log 2>&1
tail -2 gpurun_out/r2z_ncu_full.log
timeout 600 python tools/parity_report.py gpu > gpurun_out/r2z_parity.log 2>&1; tail -2 gpurun_out/r2z_parity.log
