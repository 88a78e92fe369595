#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run17.log
B="--steps 10 --warmup 3 --no-variants --no-extra --no-cpu-baseline --no-clocks"
(ESPIC_DEPOSIT_STATS=1 timeout 600 python bench.py $B --no-e2e 2> gpurun_out/r17_stats.err > gpurun_out/r17_stats.json; echo "[stats] rc=$?" > $L)
grep "deposit stats" gpurun_out/r17_stats.err | head -14 >> $L
(timeout 600 python bench.py $B 2> gpurun_out/r17_e2e.err > gpurun_out/r17_e2e.json; echo "[e2e] rc=$?" >> $L)
python -c "
import json
d=json.load(open('gpurun_out/r17_e2e.json'))
print({k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['ms_per_step'],2))
print({k:round(v,2) for k,v in d['e2e']['phases_ms'].items()}, round(d['e2e']['ms_per_step'],2))" >> $L
cat $L
