#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv1_" -s 40 -c 6 -o gpurun_out/r2m_prof_conv1 python tools/conv_lab.py --masks 1026,9218 --rounds 1 --full "" --out gpurun_out/r2m_lab.json > gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
