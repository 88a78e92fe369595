#!/bin/bash
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
echo "== smoke" > gpurun_out/run5.log
(timeout 300 python __graft_entry__.py smoke >> gpurun_out/run5.log 2>&1; echo "smoke rc=$?" >> gpurun_out/run5.log)
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multigrid or tight or mg_" 2>&1 | tail -5 >> gpurun_out/run5.log; echo "pytest rc=$?" >> gpurun_out/run5.log)
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 6 --warmup 3 $B > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "bench rc=$?" >> gpurun_out/run5.log)
grep -h "newton steps\|profile\]" gpurun_out/r5_bench.err | tail -2 >> gpurun_out/run5.log
python -c "import json; d=json.load(open('gpurun_out/r5_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['config']['pcg_iters_per_step'])" >> gpurun_out/run5.log
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 $B --mesh 256 --particles 5e7 > gpurun_out/r5_bench256.json 2> gpurun_out/r5_bench256.err; echo "bench256 rc=$?" >> gpurun_out/run5.log)
grep -h "newton steps\|profile\]" gpurun_out/r5_bench256.err | tail -2 >> gpurun_out/run5.log
cat gpurun_out/run5.log
