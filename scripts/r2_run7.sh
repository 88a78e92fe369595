#!/bin/bash
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
echo "== pytest (particle paths)" > gpurun_out/run7.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not multigrid and not tight and not mg_" 2>&1 | tail -8) >> gpurun_out/run7.log
for f in "" "--fuse" "--fuse --sort-every 16" "--fuse --fixed-point"; do
  (timeout 600 python bench.py --steps 16 --warmup 3 $B $f > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err; echo "bench [$f] rc=$?" >> gpurun_out/run7.log)
  python -c "import json; d=json.load(open('gpurun_out/r7_bench.json')); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'roofline', round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms'],2))" >> gpurun_out/run7.log
done
cat gpurun_out/run7.log
