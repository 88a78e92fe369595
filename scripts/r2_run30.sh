#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run30.log
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_shim.py -q -m gpu -k "mg_ or multigrid or solver or golden or pcg" 2>&1 | tail -3) > $L
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-variants --no-extra --no-cpu-baseline --no-clocks 2> gpurun_out/r30.err > gpurun_out/r30.json; echo "rc=$?" >> $L)
grep -h "mg profile\|mg newton profile\|profile" gpurun_out/r30.err | tail -2 >> $L
python -c "
import json
d=json.load(open('gpurun_out/r30.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'])" >> $L 2>&1
cat $L
