#!/bin/bash
# round-2 first light of the fp16 mode: GPU tests, headline bench, conv2 CTA-pair A/B, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_fp16.json 2> gpurun_out/r2a_bench_fp16.err
tail -c 600 gpurun_out/r2a_bench_fp16.json
DIM_CONV2_PAIR=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fast-mode > gpurun_out/r2a_bench_conv2pair.json 2> gpurun_out/r2a_bench_conv2pair.err
DIM_CONV_PAIR=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fast-mode > gpurun_out/r2a_bench_allpair.json 2> gpurun_out/r2a_bench_allpair.err
DIM_CONV_TAILSPLIT=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fast-mode > gpurun_out/r2a_bench_tailsplit.json 2> gpurun_out/r2a_bench_tailsplit.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --precision bf16x3 > gpurun_out/r2a_bench_bf16x3.json 2> gpurun_out/r2a_bench_bf16x3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 3 --step-batches 4 --slots 1 --no-cpu-baseline --no-fast-mode > gpurun_out/r2a_ncu_bench.log 2>&1
for f in gpurun_out/r2a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("fast_mode",{}).get("value"), d["single_stream"], d["stages_ms_per_batch_single_stream"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
