#!/bin/bash
# one B200: solver ring without the per-plane block barrier (full/empty mbarriers): parity, then timing
mkdir -p gpurun_out
L=gpurun_out/run32.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > $L
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mg_ or multigrid or golden or full_mesh" 2>&1 | tail -3) >> $L
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --no-clocks 2> gpurun_out/r32.err > gpurun_out/r32.json; echo "rc=$?" >> $L)
grep -h "mg profile" gpurun_out/r32.err | sed -n '1p;$p' >> $L
python -c "
import json
d=json.load(open('gpurun_out/r32.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'])
for k in ('config2','config5'): print(k, round(d[k]['ms_per_step'],2), round(d[k]['phases_ms']['poisson'],2), d[k]['pcg_iters_per_step'])" >> $L 2>&1
cat $L
