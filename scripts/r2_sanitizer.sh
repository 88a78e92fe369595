#!/bin/bash
# compute-sanitizer evidence for the tiny 21 x 21 x 41 case: memcheck and racecheck of smoke() -- the one-kernel Newton
# solve with its grid barriers, TMA ring and mbarriers, the tile deposit's shared-memory hash table, push, removal, injection
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log
  tail -6 gpurun_out/sanitizer_$tool.log
done
