#!/bin/bash
# one B200: grouped deposit v2 (A/B against the round-1 tile kernel), sort z-bin sweep, e2e with the k_push time, ncu of the sort scatter
mkdir -p gpurun_out
L=gpurun_out/run18.log
echo "== gpu tests" > $L
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_moments.py -q -m gpu -x 2>&1 | tail -4) >> $L
B="--steps 16 --warmup 3 --no-variants --no-extra --no-cpu-baseline --no-clocks"
run() { tag=$1; shift
  (env "$@" timeout 600 python bench.py $B 2> gpurun_out/r18_$tag.err > gpurun_out/r18_$tag.json; echo "[$tag] rc=$?" >> $L)
  python -c "
import json
d=json.load(open('gpurun_out/r18_$tag.json'))
e=d.get('e2e') or {}
print(round(d['value']/1e9,2), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'kpush', round(d['roofline']['kernel_ms'],3))
if e: print('   e2e', round(e['ms_per_step'],2), round(e['value']/d['value'],3), {k:round(v,2) for k,v in e['phases_ms'].items()}, 'kpush', round(e['k_push_ms'],3))" >> $L 2>&1
}
run new X=1
run v1 ESPIC_DEPOSIT_V1=1 --no-e2e
B="$B --no-e2e"
run zb1 ESPIC_SORT_ZBINS=1
run zb2 ESPIC_SORT_ZBINS=2
run zb4 ESPIC_SORT_ZBINS=4
N="--no-e2e --no-variants --no-cpu-baseline --no-extra --no-clocks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/r18_launches.csv python bench.py --steps 16 --warmup 3 $N --profile-range > gpurun_out/r18_launches.log 2>&1
python - >> $L <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r18_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
print('deposit/sort launches (us):', [ (r[ki][5:19], round(float(r[vi].replace(',',''))/(1e3 if r[ui]=='ns' else 1),1)) for r in rows[1:] if 'k_deposit' in r[ki] or 'k_cell' in r[ki]])
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_cell_scatter|k_deposit_group' -c 3 \
    -o gpurun_out/r18_sortdep -f python bench.py --steps 8 --warmup 3 $N --profile-range > gpurun_out/r18_ncu.log 2>&1
echo "ncu rc=$?" >> $L
cat $L
