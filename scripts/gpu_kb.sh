#!/bin/bash
# per-kernel timings + full ncu captures of the push-only and deposit kernels
mkdir -p gpurun_out
TAG=${1:-kb}
timeout 600 python scripts/kernel_bench.py --particles 2e8 2>&1 | cut -c1-200 > gpurun_out/kernel_bench_$TAG.log
cat gpurun_out/kernel_bench_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_push|k_deposit' -s 3 -c 4 \
    -o gpurun_out/pd_$TAG -f python scripts/kernel_bench.py --particles 1e8 --only "5 steps" > gpurun_out/pd_ncu_$TAG.log 2>&1
tail -3 gpurun_out/pd_ncu_$TAG.log
