#!/bin/bash
# Runs on the GPU box (under gpurun): GPU test suite, full-size bench (both arms), ncu launch list of the timed region,
# full ncu captures of the particle kernels and of the multigrid PCG kernel.  Everything lands in gpurun_out/ with the given tag.
mkdir -p gpurun_out
TAG=${1:-r1}
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
(timeout 900 python bench.py 2> gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.json)
tail -2 gpurun_out/bench_full_$TAG.err
(timeout 600 python bench.py --impl reference 2> gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.json)
tail -1 gpurun_out/bench_ref_$TAG.err
# launch list of the timed region (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --profile-range \
    > gpurun_out/launches_$TAG.log 2>&1
# the particle kernels and the Poisson kernel, full set, at the bench size
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_push|k_deposit' -c 3 \
    -o gpurun_out/particles_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --profile-range \
    > gpurun_out/particles_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg_pcg' -c 1 \
    -o gpurun_out/mg_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --profile-range \
    > gpurun_out/mg_ncu_$TAG.log 2>&1
timeout 300 python scripts/kernel_bench.py --particles 2e8 2>&1 | cut -c1-200 > gpurun_out/kernel_bench_$TAG.log
ls -la gpurun_out | tail -14
