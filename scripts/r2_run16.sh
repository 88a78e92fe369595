#!/bin/bash
# one B200: grouped deposit v2 under ncu (launch list: time per step since the sort; full capture of one launch), e2e after the DIAG/add changes
mkdir -p gpurun_out
L=gpurun_out/run16.log
echo "== gpu tests (push diag, add, deposit)" > $L
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4) >> $L
B="--steps 10 --warmup 3 --no-variants --no-extra --no-cpu-baseline --no-clocks"
run() { tag=$1; shift
  (env "$@" timeout 600 python bench.py $B 2> gpurun_out/r16_$tag.err > gpurun_out/r16_$tag.json; echo "[$tag] rc=$?" >> $L)
  python -c "
import json
d=json.load(open('gpurun_out/r16_$tag.json'))
e=d.get('e2e') or {}
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(e.get('ms_per_step',0),2), round(e.get('value',0)/d['value'],3), round(e.get('particles_per_step',0)/1e8,3))" >> $L 2>&1
}
run new X=1
run skip7 BENCH_E2E_SKIP=7
run skip1 BENCH_E2E_SKIP=1
N="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/r16_launches.csv python bench.py --steps 16 --warmup 3 $N --profile-range > gpurun_out/r16_launches.log 2>&1
echo "launch list rc=$?" >> $L
python - >> $L <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r16_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
print('deposit launches (us):', [ (r[ki][:18], round(float(r[vi].replace(',',''))/1e3 if 'ns' in rows[1] else float(r[vi].replace(',','')),1)) for r in rows[1:] if 'k_deposit' in r[ki]])
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_deposit_group' -s 4 -c 1 \
    -o gpurun_out/r16_dep -f python bench.py --steps 8 --warmup 3 $N --profile-range > gpurun_out/r16_dep_ncu.log 2>&1
echo "ncu deposit rc=$?" >> $L
cat $L
