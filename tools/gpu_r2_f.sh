#!/bin/bash
# round-2 call f: ncu --set full on the current conv / zoom / raster kernels, overfit sanity of the training step, tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_roll|conv_igemm_pair|conv_igemm_persistent|zoom_fused|raster_" -s 60 -c 20 -o gpurun_out/r2f_prof python bench.py --steps 1 --warmup 3 --step-batches 2 --slots 1 --no-cpu-baseline --no-fast-mode > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
ls -la gpurun_out/r2f_prof.ncu-rep
timeout 600 python tools/train_synth.py --overfit --steps 300 --batch 16 --lr 1e-4 --eval-every 50 --out gpurun_out/r2f_train_overfit.json > gpurun_out/r2f_train_overfit.log 2>&1
tail -8 gpurun_out/r2f_train_overfit.log
