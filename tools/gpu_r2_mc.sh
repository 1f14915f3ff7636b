#!/bin/bash
# compute-sanitizer memcheck of the final build: graph_repro (stack / roll conv1 x pair mask, eager vs graph replay, B = 4) and smoke()
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/graph_repro.py 4 > gpurun_out/r2mc_memcheck.log 2>&1
grep -v "^=========     at\|^=========     by\|^=========         in\|Host Frame\|^=========$" gpurun_out/r2mc_memcheck.log | tail -25
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2mc_memcheck_smoke.log 2>&1
tail -3 gpurun_out/r2mc_memcheck_smoke.log
