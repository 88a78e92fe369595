#!/bin/bash
# eight B200: slab solver with fewer blocks per SM (cheaper barriers?), all-reduce trace, then the full bench line at N=8
mkdir -p gpurun_out
L=gpurun_out/run23.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553"
C5="bench.py --gpus 8 --steps 6 --warmup 3 --mesh 256 --particles 1.25e8 --solver mgslab --no-extra --no-e2e --no-variants --no-clocks"
echo "== config5 variants" > $L
for v in "ESPIC_TRACE=1" "ESPIC_MG_BLOCKS_PER_SM=2" "ESPIC_MG_BLOCKS_PER_SM=1"; do
  (env $v ESPIC_MG_PROFILE=1 timeout 600 $T $C5 2> gpurun_out/r23.err > gpurun_out/r23.json; echo "[$v] rc=$?" >> $L)
  grep -h "mg slab profile" gpurun_out/r23.err | tail -1 >> $L
  grep -h "espic_deposit rank 0" gpurun_out/r23.err | tail -3 >> $L
  python -c "import json; d=json.load(open('gpurun_out/r23.json')); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'])" >> $L
done
(timeout 900 $T bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bench_n8_r2.err > gpurun_out/bench_n8_r2.json; echo "bench n8 rc=$?" >> $L)
python -c "
import json
d=json.load(open('gpurun_out/bench_n8_r2.json'))
e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e['value']/1e9,2), round(e['value']/d['value'],3))
print('parity', {k:(v.get('ok') if isinstance(v,dict) else v) for k,v in d['parity_check'].items()})
for k in ('strong','config5'): print(k, round(d[k]['value']/1e9,2), round(d[k]['ms_per_step'],2), {a:round(b,2) for a,b in d[k]['phases_ms'].items()}, d[k]['pcg_iters_per_step'])
" >> $L 2>&1
cat $L
