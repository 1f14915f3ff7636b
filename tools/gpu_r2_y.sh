#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/trained_accuracy.py --out gpurun_out/r2y_trained_accuracy.json > gpurun_out/r2y_trained.log 2>&1; tail -30 gpurun_out/r2y_trained.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q -k "flow or training_step or loss_weights" > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
