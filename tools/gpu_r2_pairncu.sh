#!/bin/bash
# per-layer durations under ncu (stable clocks, one kernel at a time): default kernel set vs the CTA-pair kernel on conv2 ... conv4_1
set -x
mkdir -p gpurun_out
for m in 8254; do
  timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"conv" -s 2720 -c 20 --csv --log-file gpurun_out/r2pn_mask$m.csv python tools/conv_lab.py --masks $m --rounds 1 --full "" --out gpurun_out/r2pn_lab$m.json > gpurun_out/r2pn_ncu$m.log 2>&1
  python - $m <<'PY'
import csv,sys
m=sys.argv[1]
rows=[r for r in csv.reader(open(f'gpurun_out/r2pn_mask{m}.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[ii],[r[ki][:40]]).append(r[vi])
for k,v in d.items(): print(m,k,v)
PY
done
