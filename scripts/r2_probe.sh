#!/bin/bash
# round-2 probe: first run of the one-kernel Newton solver.  Every step under its own timeout (a grid-barrier bug would hang).
mkdir -p gpurun_out
echo "== smoke" > gpurun_out/probe.log
(timeout 300 python __graft_entry__.py smoke >> gpurun_out/probe.log 2>&1; echo "smoke rc=$?" >> gpurun_out/probe.log)
echo "== solver tests" >> gpurun_out/probe.log
(ESPIC_MG_PROFILE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multigrid or tight or properties_at_baseline or golden" >> gpurun_out/probe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/probe.log)
echo "== bench (no extras), profile" >> gpurun_out/probe.log
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --no-extra --dump-warm gpurun_out/warm_128.npz \
   > gpurun_out/probe_default.json 2> gpurun_out/probe_default.err; echo "bench rc=$?" >> gpurun_out/probe.log)
(ESPIC_MG_PROFILE=1 ESPIC_MG_EXACT_NEWTON=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --no-extra \
   > gpurun_out/probe_exact.json 2> gpurun_out/probe_exact.err; echo "bench exact rc=$?" >> gpurun_out/probe.log)
grep -h "newton steps\|profile\]" gpurun_out/probe_default.err | tail -12 >> gpurun_out/probe.log
tail -40 gpurun_out/probe.log
cut -c1-900 gpurun_out/probe_default.json
