#!/bin/bash
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
echo "== smoke" > gpurun_out/run4.log
(timeout 300 python __graft_entry__.py smoke >> gpurun_out/run4.log 2>&1; echo "smoke rc=$?" >> gpurun_out/run4.log)
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_shim.py -x -q -m gpu -k "multigrid or tight or mg_ or vti or properties" 2>&1 | tail -15 >> gpurun_out/run4.log; echo "pytest rc=$?" >> gpurun_out/run4.log)
for cfg in "0.6 1.5" "0.5 1.5" "0.7 1.5" "0.6 2.0" "0.6 1.25" "1.0 2.0"; do
  set -- $cfg
  (ESPIC_MG_LINK_SCALE=$1 ESPIC_MG_ETA_POW=$2 ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 6 --warmup 3 $B > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench_$1_$2.err; echo "scale $1 pow $2 bench rc=$?" >> gpurun_out/run4.log)
  grep -h "newton steps" gpurun_out/r4_bench_$1_$2.err | tail -2 >> gpurun_out/run4.log
  python -c "import json; d=json.load(open('gpurun_out/r4_bench.json')); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'], d['config']['newton_iters_per_step'])" >> gpurun_out/run4.log
done
cat gpurun_out/run4.log
