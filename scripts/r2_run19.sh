#!/bin/bash
mkdir -p gpurun_out
N="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_cell_scatter|k_cell_count' -c 2 \
    -o gpurun_out/r19_sort -f python bench.py --steps 8 --warmup 3 $N --profile-range > gpurun_out/r19_ncu.log 2>&1
echo "ncu rc=$?"
