#!/bin/bash
# eight B200, final build: the full bench line
mkdir -p gpurun_out
L=gpurun_out/run35.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556"
(timeout 900 $T bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bench_n8_final.err > gpurun_out/bench_n8_final.json; echo "bench n8 rc=$?" > $L)
python -c "
import json
d=json.load(open('gpurun_out/bench_n8_final.json'))
e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e['value']/1e9,2), round(e['value']/d['value'],3))
print('parity', {k:(v.get('ok') if isinstance(v,dict) else v) for k,v in d['parity_check'].items()})
for k in ('strong','config5'): print(k, round(d[k]['value']/1e9,2), round(d[k]['ms_per_step'],2), {a:round(b,2) for a,b in d[k]['phases_ms'].items()}, d[k]['pcg_iters_per_step'])
" >> $L 2>&1
cat $L
