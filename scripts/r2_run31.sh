#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run31.log
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_moments.py -q -m gpu 2>&1 | tail -3) > $L
(timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-variants --no-extra --no-cpu-baseline --no-clocks 2> gpurun_out/r31.err > gpurun_out/r31.json; echo "rc=$?" >> $L)
python -c "
import json
d=json.load(open('gpurun_out/r31.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'])" >> $L 2>&1
N="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/r31_launches.csv python bench.py --steps 8 --warmup 3 $N --profile-range > gpurun_out/r31_launches.log 2>&1
python - >> $L <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r31_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
print('deposit launches (us):', [ round(float(r[vi].replace(',',''))/(1e3 if r[ui]=='ns' else 1),1) for r in rows[1:] if 'k_deposit' in r[ki]])
PY
cat $L
