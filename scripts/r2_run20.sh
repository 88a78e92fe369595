#!/bin/bash
# one B200: gather-based sort + deposit with per-run REDs: parity tests, bench, launch list
mkdir -p gpurun_out
L=gpurun_out/run20.log
echo "== gpu tests" > $L
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_moments.py tests/test_migration.py -q -m gpu -x 2>&1 | tail -4) >> $L
B="--steps 16 --warmup 3 --no-variants --no-extra --no-cpu-baseline --no-clocks"
run() { tag=$1; shift
  (env "$@" timeout 600 python bench.py $B 2> gpurun_out/r20_$tag.err > gpurun_out/r20_$tag.json; echo "[$tag] rc=$?" >> $L)
  python -c "
import json
d=json.load(open('gpurun_out/r20_$tag.json'))
e=d.get('e2e') or {}
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'kpush', round(d['roofline']['kernel_ms'],3))
if e: print('   e2e', round(e['ms_per_step'],2), round(e['value']/d['value'],3), {k:round(v,2) for k,v in e['phases_ms'].items()}, 'kpush', round(e['k_push_ms'],3))" >> $L 2>&1
}
run new X=1
B="$B --no-e2e"
run zb2 ESPIC_SORT_ZBINS=2
run zb4 ESPIC_SORT_ZBINS=4
N="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/r20_launches.csv python bench.py --steps 16 --warmup 3 $N --profile-range > gpurun_out/r20_launches.log 2>&1
python - >> $L <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r20_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
print('deposit/sort launches (us):', [ (r[ki].replace('void ','')[:14], round(float(r[vi].replace(',',''))/(1e3 if r[ui]=='ns' else 1),1)) for r in rows[1:] if 'k_deposit' in r[ki] or 'k_cell' in r[ki]])
PY
cat $L
