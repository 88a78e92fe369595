#!/bin/bash
# one B200, final build: ncu launch list and full captures (push, grouped deposit, sort passes, solver)
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 8 --warmup 3 $B --profile-range > gpurun_out/r2_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_push|k_deposit_group' -c 2 \
    -o gpurun_out/r2_particles -f python bench.py --steps 2 --warmup 3 $B --profile-range > gpurun_out/r2_particles_ncu.log 2>&1
echo "ncu particles rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg_newton' -c 1 \
    -o gpurun_out/r2_mg_newton -f python bench.py --steps 1 --warmup 3 $B --profile-range > gpurun_out/r2_mg_ncu.log 2>&1
echo "ncu solver rc=$?"
