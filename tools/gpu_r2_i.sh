#!/bin/bash
set -x
mkdir -p gpurun_out
for T in 900 1800; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fast-mode --train-steps $T > gpurun_out/r2i_bench_t$T.json 2> gpurun_out/r2i_bench_t$T.err
python - gpurun_out/r2i_bench_t$T.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], json.dumps(d["add_m"])[:700])
PY
done
