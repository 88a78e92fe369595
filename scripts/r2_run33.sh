#!/bin/bash
# one B200, final build of the round: sanitizer on smoke(), complete GPU suite, bench (20/5) and reference arm
mkdir -p gpurun_out
L=gpurun_out/run33.log
bash scripts/r2_sanitizer.sh > $L 2>&1
echo "== pytest -m gpu (all)" >> $L
(timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4) >> $L
(timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_n1_final.err > gpurun_out/bench_n1_final.json; echo "bench rc=$?" >> $L)
(timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/bench_ref_final.err > gpurun_out/bench_ref_final.json; echo "reference arm rc=$?" >> $L)
python -c "
import json
d=json.load(open('gpurun_out/bench_n1_final.json')); e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e['ms_per_step'],2), round(e['value']/d['value'],3), 'frac', round(d['roofline']['frac'],3), round(d['roofline_poisson']['frac'],3))
for k in ('config2','config5'): print(k, round(d[k]['value']/1e9,2), round(d[k]['ms_per_step'],2), {a:round(b,2) for a,b in d[k]['phases_ms'].items()})
r=json.load(open('gpurun_out/bench_ref_final.json')); print('reference', r['value'], r['ms_per_step'])
" >> $L 2>&1
cat $L
