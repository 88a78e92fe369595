#!/bin/bash
# two B200, final state: complete GPU suite (the multi-GPU tests included), default bench at N=1 and N=2
mkdir -p gpurun_out
L=gpurun_out/run29.log
echo "== pytest -m gpu (all, 2 GPUs visible)" > $L
(timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -5) >> $L
(timeout 900 python bench.py 2> gpurun_out/r29_n1.err > gpurun_out/r29_n1.json; echo "bench n1 (defaults) rc=$?" >> $L)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> $L
python -c "
import json
d=json.load(open('gpurun_out/r29_n1.json')); e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e['value']/d['value'],3), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])" >> $L 2>&1
cat $L
