#!/bin/bash
# one B200: coarse-level damping 1.0 and more sweeps on the coarsest level: solver parity tests, then iterations and time
mkdir -p gpurun_out
L=gpurun_out/run36.log
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mg_ or multigrid or golden or full_mesh or solver" 2>&1 | tail -2) > $L
B="--steps 16 --warmup 3 --no-e2e --no-variants --no-extra --no-cpu-baseline --no-clocks"
for cs in 6 9 13; do
  (ESPIC_MG_COARSE_SWEEPS=$cs ESPIC_MG_PROFILE=1 timeout 600 python bench.py $B 2> gpurun_out/r36.err > gpurun_out/r36.json; echo "[coarsest sweeps 1+$cs] rc=$?" >> $L)
  grep -h "mg profile" gpurun_out/r36.err | tail -1 >> $L
  python -c "
import json
d=json.load(open('gpurun_out/r36.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), 'poisson', round(d['phases_ms']['poisson'],3), d['config']['pcg_iters_per_step'], d['config']['newton_iters_per_step'])" >> $L 2>&1
done
cat $L
