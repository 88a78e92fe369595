#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
C5="bench.py --gpus $N --steps 6 --warmup 3 --mesh 256 --particles 2e7 --solver mgslab --no-extra --no-e2e --no-variants --no-clocks"
echo "== slab blocks per SM, N=$N" > gpurun_out/run13.log
for v in 3 2 1; do
  (ESPIC_MG_BLOCKS_PER_SM=$v ESPIC_MG_PROFILE=1 timeout 600 $T $C5 2> gpurun_out/r13.err > gpurun_out/r13.json; echo "[bps $v] rc=$?" >> gpurun_out/run13.log)
  grep -h "mg slab profile" gpurun_out/r13.err | tail -1 >> gpurun_out/run13.log
  python -c "import json; d=json.load(open('gpurun_out/r13.json')); print(round(d['phases_ms']['poisson'],2), d['config']['pcg_iters_per_step'])" >> gpurun_out/run13.log
done
echo "== single GPU 128^3 blocks per SM" >> gpurun_out/run13.log
for v in 3 2 1; do
  (ESPIC_MG_BLOCKS_PER_SM=$v ESPIC_MG_PROFILE=1 timeout 600 python bench.py --steps 6 --warmup 3 --particles 2e7 --no-extra --no-e2e --no-variants --no-clocks --no-cpu-baseline 2> gpurun_out/r13.err > gpurun_out/r13.json; echo "[bps $v] rc=$?" >> gpurun_out/run13.log)
  grep -h "mg profile" gpurun_out/r13.err | tail -1 >> gpurun_out/run13.log
  python -c "import json; d=json.load(open('gpurun_out/r13.json')); print(round(d['phases_ms']['poisson'],2), d['config']['pcg_iters_per_step'])" >> gpurun_out/run13.log
done
cat gpurun_out/run13.log
