#!/bin/bash
# 2 GPUs: multi-GPU parity tests, then the bench with parity_check / strong / config5 records
mkdir -p gpurun_out
echo "== pytest multigpu" > gpurun_out/run8.log
(timeout 1200 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -25) >> gpurun_out/run8.log
echo "== bench N=2" >> gpurun_out/run8.log
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 \
   2> gpurun_out/bench_n2_r2.err > gpurun_out/bench_n2_r2.json; echo "bench rc=$?" >> gpurun_out/run8.log)
grep -h "parity check\|strong\|config5\|timed region\|Error\|error" gpurun_out/bench_n2_r2.err | cut -c1-900 | head -20 >> gpurun_out/run8.log
cat gpurun_out/run8.log
