#!/bin/bash
# eight B200: does the host-side part of the e2e step shrink when waiting host threads block instead of spinning?
mkdir -p gpurun_out
L=gpurun_out/run28.log
nproc > $L
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555"
for v in default block; do
  (ESPIC_SYNC_MODE=$v timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --no-variants --no-cpu-baseline --no-clocks 2> gpurun_out/r28_$v.err > gpurun_out/r28_$v.json; echo "[$v] rc=$?" >> $L)
  python -c "
import json
d=json.load(open('gpurun_out/r28_$v.json')); e=d['e2e']
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), 'e2e', round(e['ms_per_step'],2), round(e['value']/d['value'],3), {k:round(v,2) for k,v in e['phases_ms'].items()})" >> $L 2>&1
done
cat $L
