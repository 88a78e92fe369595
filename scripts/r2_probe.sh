#!/bin/bash
# round-2 probe: Newton/CG residual histories of the shipped solver at the bench size, and the warm Poisson problem of one step
mkdir -p gpurun_out
ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --dump-warm gpurun_out/warm_128.npz \
   > gpurun_out/probe_default.json 2> gpurun_out/probe_default.err
ESPIC_MG_PROFILE=1 ESPIC_MG_EXACT_NEWTON=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-variants --no-cpu-baseline \
   > gpurun_out/probe_exact.json 2> gpurun_out/probe_exact.err
grep -c newton gpurun_out/probe_default.err
tail -3 gpurun_out/probe_default.json | cut -c1-600
