#!/bin/bash
# one B200, final build of the round: complete GPU suite and the bench line
mkdir -p gpurun_out
L=gpurun_out/run37.log
echo "== pytest -m gpu (all)" > $L
(timeout 240 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) >> $L
(timeout 120 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_n1_final2.err > gpurun_out/bench_n1_final2.json; echo "bench rc=$?" >> $L)
python -c "
import json
d=json.load(open('gpurun_out/bench_n1_final2.json')); e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e['ms_per_step'],2), round(e['value']/d['value'],3), 'frac', round(d['roofline']['frac'],3), round(d['roofline_poisson']['frac'],3), d['config']['pcg_iters_per_step'])
for k in ('config2','config5'): print(k, round(d[k]['value']/1e9,2), round(d[k]['ms_per_step'],2), {a:round(b,2) for a,b in d[k]['phases_ms'].items()}, d[k]['pcg_iters_per_step'])
" >> $L 2>&1
cat $L
