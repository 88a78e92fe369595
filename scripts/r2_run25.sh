#!/bin/bash
# one B200: sort interval with the three-pass sort
mkdir -p gpurun_out
L=gpurun_out/run25.log
: > $L
for se in 8 12 16 24; do
  (timeout 600 python bench.py --steps 48 --warmup 3 --sort-every $se --no-e2e --no-variants --no-extra --no-cpu-baseline --no-clocks 2> gpurun_out/r25.err > gpurun_out/r25.json; echo "[sort every $se] rc=$?" >> $L)
  python -c "
import json
d=json.load(open('gpurun_out/r25.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'kpush', round(d['roofline']['kernel_ms'],3))" >> $L 2>&1
done
cat $L
