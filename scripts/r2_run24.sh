#!/bin/bash
# round-2 evidence run on one B200: full GPU test suite, sanitizer logs, ncu launch list + full captures, full bench + reference arm
mkdir -p gpurun_out
B="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
echo "== pytest -m gpu (all)" > gpurun_out/run24.log
(timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -12) >> gpurun_out/run24.log
bash scripts/r2_sanitizer.sh >> gpurun_out/run24.log 2>&1
echo "== ncu launch list" >> gpurun_out/run24.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 8 --warmup 3 $B --profile-range > gpurun_out/r2_launches.log 2>&1
echo "launch list rc=$?" >> gpurun_out/run24.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_push|k_deposit_group' -c 2 \
    -o gpurun_out/r2_particles -f python bench.py --steps 2 --warmup 3 $B --profile-range > gpurun_out/r2_particles_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_cell_' -c 3 \
    -o gpurun_out/r2_sort -f python bench.py --steps 8 --warmup 3 $B --profile-range > gpurun_out/r2_sort_ncu.log 2>&1
echo "ncu particles rc=$?" >> gpurun_out/run24.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg_newton' -c 1 \
    -o gpurun_out/r2_mg_newton -f python bench.py --steps 1 --warmup 3 $B --profile-range > gpurun_out/r2_mg_ncu.log 2>&1
echo "ncu solver rc=$?" >> gpurun_out/run24.log
echo "== full bench" >> gpurun_out/run24.log
(timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_n1_r2.err > gpurun_out/bench_n1_r2.json; echo "bench rc=$?" >> gpurun_out/run24.log)
(timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/bench_ref_r2.err > gpurun_out/bench_ref_r2.json; echo "reference arm rc=$?" >> gpurun_out/run24.log)
python -c "
import json
d=json.load(open('gpurun_out/bench_n1_r2.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/d['value'],3), 'frac', round(d['roofline']['frac'],3), round(d['roofline_poisson']['frac'],3))
for k in ('config2','config5'): print(k, round(d[k]['value']/1e9,2), round(d[k]['ms_per_step'],2), {a:round(b,2) for a,b in d[k]['phases_ms'].items()})
r=json.load(open('gpurun_out/bench_ref_r2.json')); print('reference', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'])
" >> gpurun_out/run24.log 2>&1
cat gpurun_out/run24.log
