#!/bin/bash
# 8 GPUs: the full bench line (headline weak scaling + parity_check + strong + config5), then config5 alone with the solver profile
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
(timeout 1500 $T bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bench_n8_r2.err > gpurun_out/bench_n8_r2.json; echo "bench rc=$?" > gpurun_out/run11.log)
grep -h "parity check\|strong:\|config5:\|timed region" gpurun_out/bench_n8_r2.err | sort -u | cut -c1-1100 | head -8 >> gpurun_out/run11.log
(ESPIC_MG_PROFILE=1 timeout 900 $T bench.py --gpus 8 --steps 6 --warmup 3 --mesh 256 --particles 1.25e8 --solver mgslab --no-extra --no-e2e --no-variants --no-clocks \
   2> gpurun_out/bench_n8_c5prof.err > gpurun_out/bench_n8_c5prof.json; echo "config5 profile rc=$?" >> gpurun_out/run11.log)
grep -h "mg slab" gpurun_out/bench_n8_c5prof.err | tail -4 >> gpurun_out/run11.log
python -c "import json; d=json.load(open('gpurun_out/bench_n8_c5prof.json')); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'])" >> gpurun_out/run11.log
cat gpurun_out/run11.log
