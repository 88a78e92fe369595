#!/bin/bash
# 8 GPUs: slab-solver variants on configs[4] (256^3, 1.25e8 ions per GPU), solver profile on
mkdir -p gpurun_out
echo "== multigpu tests (2 of the 8 GPUs)" > gpurun_out/run12.log
(timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -4) >> gpurun_out/run12.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
C5="bench.py --gpus 8 --steps 6 --warmup 3 --mesh 256 --particles 1.25e8 --solver mgslab --no-extra --no-e2e --no-variants --no-clocks"
for v in "ESPIC_MG_SLAB_NEIGHBOUR_SYNC=1" "ESPIC_MG_SLAB_NEIGHBOUR_SYNC=0" "ESPIC_MG_SLAB_REDUNDANT_NODES=600000" "ESPIC_MG_SLAB_REDUNDANT_NODES=8192"; do
  (env $v ESPIC_MG_PROFILE=1 timeout 600 $T $C5 2> gpurun_out/r12.err > gpurun_out/r12.json; echo "[$v] rc=$?" >> gpurun_out/run12.log)
  grep -h "mg slab profile" gpurun_out/r12.err | tail -1 >> gpurun_out/run12.log
  python -c "import json; d=json.load(open('gpurun_out/r12.json')); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['pcg_iters_per_step'])" >> gpurun_out/run12.log
done
cat gpurun_out/run12.log
