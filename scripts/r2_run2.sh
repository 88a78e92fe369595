#!/bin/bash
# round-2 second GPU run: ncu of the Newton kernel, full GPU test suite, sort-order sweep, full bench with the extra records
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
echo "== ncu k_mg_newton" > gpurun_out/run2.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_mg_newton -c 1 \
    -o gpurun_out/mgn_r2a -f python bench.py --steps 1 --warmup 3 $B --profile-range > gpurun_out/mgn_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/run2.log
echo "== sort sweep" >> gpurun_out/run2.log
for cfg in "xtoc 8 0" "drift 8 12" "drift 16 12" "drift 16 100" "drift 32 100" "drift 32 0"; do
  set -- $cfg
  ESPIC_DEPOSIT_PLAIN_STEPS=$3 timeout 300 python bench.py --steps 32 --warmup 3 $B --sort-order $1 --sort-every $2 2> gpurun_out/sweep_$1_$2_$3.err \
     | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],3))" >> gpurun_out/run2.log 2>&1
done
echo "== pytest -m gpu" >> gpurun_out/run2.log
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25) >> gpurun_out/run2.log
echo "== full bench" >> gpurun_out/run2.log
(timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_full_r2a.err > gpurun_out/bench_full_r2a.json; echo "bench rc=$?" >> gpurun_out/run2.log)
tail -5 gpurun_out/bench_full_r2a.err >> gpurun_out/run2.log
cat gpurun_out/run2.log
