#!/bin/bash
mkdir -p gpurun_out
echo "== new tests" > gpurun_out/run9.log
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_shim.py tests/test_migration.py -x -q -m gpu -k "diag or prefetch or two_species or class_api or ch3_main or ch9 or vti" 2>&1 | tail -8) >> gpurun_out/run9.log
echo "== bench (headline + e2e, no extras)" >> gpurun_out/run9.log
(timeout 900 python bench.py --steps 10 --warmup 3 --no-extra 2> gpurun_out/r9_bench.err > gpurun_out/r9_bench.json; echo "bench rc=$?" >> gpurun_out/run9.log)
python -c "import json; d=json.load(open('gpurun_out/r9_bench.json')); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/d['value'],3), 'qn', round(d['solver_variants']['qn']['ms_per_step'],2))" >> gpurun_out/run9.log
echo "== reference arm" >> gpurun_out/run9.log
(time timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2> gpurun_out/r9_ref.err > gpurun_out/r9_ref.json) 2>> gpurun_out/run9.log
python -c "import json; d=json.load(open('gpurun_out/r9_ref.json')); print(d['value'], d['ms_per_step'], d['cpu_baseline']['phases_s'], d['cpu_baseline']['cores'], d['solver_variants'])" >> gpurun_out/run9.log
cat gpurun_out/run9.log
