#!/bin/bash
# one B200: new grouped deposit kernel (parity tests + A/B against the round-1 tile kernel), e2e ablations
mkdir -p gpurun_out
L=gpurun_out/run15.log
echo "== deposit-related gpu tests" > $L
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_moments.py -q -m gpu -x 2>&1 | tail -6) >> $L
B="--steps 10 --warmup 3 --no-variants --no-extra --no-cpu-baseline --no-clocks"
run() { # tag, env...
  tag=$1; shift
  (env "$@" timeout 600 python bench.py $B 2> gpurun_out/r15_$tag.err > gpurun_out/r15_$tag.json; echo "[$tag] rc=$?" >> $L)
  python -c "
import json
d=json.load(open('gpurun_out/r15_$tag.json'))
e=d.get('e2e') or {}
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e.get('ms_per_step',0),2), round(e.get('value',0)/d['value'],3), round(e.get('particles_per_step',0)/1e8,3))" >> $L 2>&1
}
run new X=1
run v1 ESPIC_DEPOSIT_V1=1
run skip7 BENCH_E2E_SKIP=7
run skip1 BENCH_E2E_SKIP=1
run skip2 BENCH_E2E_SKIP=2
run skip4 BENCH_E2E_SKIP=4

cat $L
