#!/bin/bash
# one B200: full GPU suite after the particle-phase changes, bench with extras
mkdir -p gpurun_out
L=gpurun_out/run21.log
echo "== pytest -m gpu (all)" > $L
(timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6) >> $L
(timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r21_bench.err > gpurun_out/r21_bench.json; echo "bench rc=$?" >> $L)
python -c "
import json
d=json.load(open('gpurun_out/r21_bench.json'))
e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'kpush', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))
print('   e2e', round(e['ms_per_step'],2), round(e['value']/d['value'],3), {k:round(v,2) for k,v in e['phases_ms'].items()}, 'kpush', round(e['k_push_ms'],3))
for k in ('config2','config5'): print(k, round(d[k]['value']/1e9,2), round(d[k]['ms_per_step'],2), {a:round(b,2) for a,b in d[k]['phases_ms'].items()})
print('qn', d['solver_variants']['qn']['value'], 'cpu', d['cpu_baseline']['value'])
" >> $L 2>&1
cat $L
