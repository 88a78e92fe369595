#!/bin/bash
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
echo "== smoke" > gpurun_out/run3.log
(timeout 300 python __graft_entry__.py smoke >> gpurun_out/run3.log 2>&1; echo "smoke rc=$?" >> gpurun_out/run3.log)
(ESPIC_MG_PROFILE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_shim.py -x -q -m gpu -k "multigrid or tight or mg_ or golden or vti or properties" 2>&1 | tail -30 >> gpurun_out/run3.log; echo "pytest rc=$?" >> gpurun_out/run3.log)
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 6 --warmup 3 $B > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?" >> gpurun_out/run3.log)
grep -h "newton steps\|profile\]" gpurun_out/r3_bench.err | tail -4 >> gpurun_out/run3.log
python -c "import json; d=json.load(open('gpurun_out/r3_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['config']['pcg_iters_per_step'])" >> gpurun_out/run3.log
(ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 $B --mesh 256 --particles 5e7 > gpurun_out/r3_bench256.json 2> gpurun_out/r3_bench256.err; echo "bench256 rc=$?" >> gpurun_out/run3.log)
grep -h "newton steps\|profile\]" gpurun_out/r3_bench256.err | tail -2 >> gpurun_out/run3.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_mg_newton -c 1 \
    -o gpurun_out/mgn_r2b -f python bench.py --steps 1 --warmup 3 $B --profile-range > gpurun_out/mgn_ncu_b.log 2>&1
echo "ncu rc=$?" >> gpurun_out/run3.log
cat gpurun_out/run3.log
