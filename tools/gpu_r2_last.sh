#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2last_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2last_smoke.log 2>&1; tail -1 gpurun_out/r2last_smoke.log
