#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (all)" > gpurun_out/run6.log
(timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25) >> gpurun_out/run6.log
echo "== full bench" >> gpurun_out/run6.log
(timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_full_r2b.err > gpurun_out/bench_full_r2b.json; echo "bench rc=$?" >> gpurun_out/run6.log)
grep -h "config2\|config5\|timed region" gpurun_out/bench_full_r2b.err | cut -c1-700 >> gpurun_out/run6.log
cat gpurun_out/run6.log
