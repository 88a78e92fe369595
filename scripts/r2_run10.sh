#!/bin/bash
mkdir -p gpurun_out
echo "== e2e modes" > gpurun_out/run10.log
for m in 0 1 2 4 7; do
(timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-variants --no-cpu-baseline --no-clocks --e2e-mode $m 2> gpurun_out/r10_bench.err > gpurun_out/r10_bench.json; echo "mode $m rc=$?" >> gpurun_out/run10.log)
python -c "import json; d=json.load(open('gpurun_out/r10_bench.json')); print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/d['value'],3))" >> gpurun_out/run10.log
done
cat gpurun_out/run10.log
