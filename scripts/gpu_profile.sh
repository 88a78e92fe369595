#!/bin/bash
# Runs on the GPU box (under gpurun): full-size bench, ncu launch list of the timed region, one full capture of the push kernel.
mkdir -p gpurun_out
TAG=${1:-r1}
(timeout 900 python bench.py 2> gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.json) 
tail -3 gpurun_out/bench_full_$TAG.err
# launch list of the timed region (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --particles 5e7 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profile-range \
    > gpurun_out/launches_$TAG.log 2>&1
# the dominant kernel, full set
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_push -c 1 \
    -o gpurun_out/push_$TAG -f python bench.py --particles 5e7 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profile-range \
    > gpurun_out/push_ncu_$TAG.log 2>&1
ls -la gpurun_out
